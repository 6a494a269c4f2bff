#!/bin/bash
# round 2, second GPU call: full GPU test suite with the round-2 tests, item-schedule tuning on a strip-shaped grid,
# the parity tables with per-cell counts, a bench line with the merged tail launch
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02b_pytest.txt
tail -4 gpurun_out/r02b_pytest.txt
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f launches %d' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['gpu_launches']))"; }
{
run "A=1" ""
for q in "1,1,1" "0.5,0.5,1" "0,1,1" "0.5,1,2" "0,0.5,2" "0,0,2" "1,1,2"; do run "SWE_B200_X2_Q=$q" "--ny 2048"; done
run "SWE_B200_X2_ROWS=63" "--ny 2048"
run "SWE_B200_X2_ROWS=90" "--ny 2048"
} > gpurun_out/r02b_ab.txt 2>&1
cat gpurun_out/r02b_ab.txt
timeout 900 python scratch/parity_report.py --out gpurun_out/r02b_parity.md > /dev/null 2> gpurun_out/r02b_parity.err
tail -3 gpurun_out/r02b_parity.err
