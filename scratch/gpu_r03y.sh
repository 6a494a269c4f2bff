#!/bin/bash
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 6 -c 2 -f -o gpurun_out/r03y_x2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03y_ncu_x2.log 2>&1
tail -2 gpurun_out/r03y_ncu_x2.log; ls -la gpurun_out/r03y_x2.ncu-rep
