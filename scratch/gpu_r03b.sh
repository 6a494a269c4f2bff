#!/bin/bash
# instruction diet of the x2 kernels: single-instruction sub.sat (X2_SUBSAT), all-dry warp-rows without the x reconstruction of the
# momenta and the flux differences (X2_DRYLEAN); bitwise check against the old build, then interleaved A/B
mkdir -p gpurun_out
V=$PWD/scratch/variants
for v in old lean; do SWE_B200_LIB=$V/libswe_b200_$v.so timeout 120 python scratch/dump_state.py gpurun_out/r03b_state_$v.npz 2> gpurun_out/r03b_dump_$v.err; done
python scratch/dump_state.py --compare gpurun_out/r03b_state_old.npz gpurun_out/r03b_state_lean.npz > gpurun_out/r03b_bitwise.txt 2>&1; cat gpurun_out/r03b_bitwise.txt
rm -f gpurun_out/r03b_state_*.npz
run() { echo -n "$1 $3: "; env $2 timeout 200 python bench.py --no-cpu --steps 20 --warmup 5 $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run old "SWE_B200_LIB=$V/libswe_b200_old.so" ""
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" ""
run lean "SWE_B200_LIB=$V/libswe_b200_lean.so" ""
run leanng "SWE_B200_LIB=$V/libswe_b200_leanng.so" ""
run old "SWE_B200_LIB=$V/libswe_b200_old.so" ""
run lean "SWE_B200_LIB=$V/libswe_b200_lean.so" ""
run old "SWE_B200_LIB=$V/libswe_b200_old.so" "--level 60"
run sat "SWE_B200_LIB=$V/libswe_b200_sat.so" "--level 60"
run lean "SWE_B200_LIB=$V/libswe_b200_lean.so" "--level 60"
} > gpurun_out/r03b_ab.txt 2>&1
cat gpurun_out/r03b_ab.txt
