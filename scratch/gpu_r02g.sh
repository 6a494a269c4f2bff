#!/bin/bash
# round 2: NeumannBC on the fused path (tests + timings of the reference examples' BC), occupancy sensitivity of the x2 kernels
mkdir -p gpurun_out
V=$PWD/scratch/variants
( timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_randomized.py tests/test_gpu_parity.py -m gpu -q --maxfail=12 2>&1 | tail -40 ) > gpurun_out/r02g_pytest.txt
tail -6 gpurun_out/r02g_pytest.txt
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f launches %d clocks %s %s' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run "A=1" ""
run "SWE_B200_LIB=$V/libswe_b200_occ3.so" ""
run "SWE_B200_LIB=$V/libswe_b200_occ2.so" ""
run "A=1" "--config 3 --bc neumann"
run "A=1" "--config 3 --bc neumann --unfused"
run "SWE_B200_NO_X2=1" "--config 3"
run "A=1" "--config 4 --bc neumann"
run "A=1" "--config 4 --bc neumann --unfused"
run "SWE_B200_NO_X2=1" "--config 4"
run "A=1" "--config 3 --bc neumann --dtype float64"
run "A=1" "--config 4 --bc neumann --dtype float64 --n 4096"
} > gpurun_out/r02g_ab.txt 2>&1
cat gpurun_out/r02g_ab.txt
