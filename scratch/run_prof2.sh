#!/bin/bash
mkdir -p gpurun_out
V=$PWD/scratch/variants
for v in d n; do
SWE_B200_LIB=$V/libswe_b200_$v.so ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 7 -c 1 -o gpurun_out/p2_$v -f python bench.py --n 8192 --steps 3 --warmup 3 --no-cpu > gpurun_out/p2_$v.log 2>&1
done
export BENCH_ARGS="--n 8192"
scratch/ab.sh "SWE_B200_LIB=$V/libswe_b200_d.so" "SWE_B200_LIB=$V/libswe_b200_n.so" | tee gpurun_out/ab4.txt
