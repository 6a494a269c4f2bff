#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02m_pytest.txt
tail -4 gpurun_out/r02m_pytest.txt
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f launches %d clocks %s %s' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run "A=1" ""
run "A=1" ""
run "A=1" "--config 3"
run "A=1" "--config 4"
run "SWE_B200_NO_D1=1" "--config 2 --n 4096"
} > gpurun_out/r02m_ab.txt 2>&1
cat gpurun_out/r02m_ab.txt
