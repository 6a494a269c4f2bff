#!/bin/bash
mkdir -p gpurun_out
V=$PWD/scratch/variants
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac']))"; }
{
for i in 1 2; do
run "A=1" "--config 2 --size 4096"
run "SWE_B200_LIB=$V/libswe_b200_d1mm.so" "--config 2 --size 4096"
run "A=1" "--dtype float64 --size 8192"
run "SWE_B200_LIB=$V/libswe_b200_d1mm.so" "--dtype float64 --size 8192"
done
} > gpurun_out/r02v_ab.txt 2>&1
cat gpurun_out/r02v_ab.txt
