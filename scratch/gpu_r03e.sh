#!/bin/bash
mkdir -p gpurun_out
V=$PWD/scratch/variants
run() { echo -n "$1 $3: "; env $2 timeout 30 python bench.py --no-cpu --steps 20 --warmup 5 $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
run ng2 "SWE_B200_LIB=$V/libswe_b200_ng2.so" "--size 4096" > gpurun_out/r03e_ng2.txt 2>&1 &
wait
run ship "A=1" "--size 4096" > gpurun_out/r03e_ship.txt 2>&1
cat gpurun_out/r03e_ng2.txt gpurun_out/r03e_ship.txt
