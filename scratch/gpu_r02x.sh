#!/bin/bash
# re-check of round 1's compile-time tunings on the final kernels (interleaved with the default)
mkdir -p gpurun_out
V=$PWD/scratch/variants
run() { echo -n "$1 : "; env $2 python bench.py --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
for v in cu3 cu0 tc3 tc0; do
run "default" "A=1"
run "$v" "SWE_B200_LIB=$V/libswe_b200_$v.so"
done
run "default" "A=1"
} > gpurun_out/r02x_ab.txt 2>&1
cat gpurun_out/r02x_ab.txt
