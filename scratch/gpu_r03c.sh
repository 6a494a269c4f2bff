#!/bin/bash
# X2_NOGHOST_FAST re-test on the sub.sat build (round 1 measured it +5 % on the kernels of that time)
mkdir -p gpurun_out
V=$PWD/scratch/variants
run() { echo -n "$1 $3: "; env $2 timeout 200 python bench.py --no-cpu --steps 20 --warmup 5 $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" ""
run satng "SWE_B200_LIB=$V/libswe_b200_satng.so" ""
run leanng "SWE_B200_LIB=$V/libswe_b200_leanng.so" ""
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" ""
run satng "SWE_B200_LIB=$V/libswe_b200_satng.so" ""
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" "--level 60"
run satng "SWE_B200_LIB=$V/libswe_b200_satng.so" "--level 60"
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" "--config 3"
run satng "SWE_B200_LIB=$V/libswe_b200_satng.so" "--config 3"
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" "--config 4"
run satng "SWE_B200_LIB=$V/libswe_b200_satng.so" "--config 4"
} > gpurun_out/r03c_ab.txt 2>&1
cat gpurun_out/r03c_ab.txt
