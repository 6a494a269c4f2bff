#!/bin/bash
# round 2: one-wave item schedule for tiny grids (A/B), full GPU suite on the result
mkdir -p gpurun_out
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f launches %d' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['gpu_launches']))"; }
{
for a in "--config 2" "--config 2 --n 2048" "--config 3 --n 1024" "--config 3 --n 2048" "--config 4 --n 1024" "--config 5 --n 1024" "--config 5 --n 2048" "--config 2 --n 512"; do
run "A=1" "$a"
run "SWE_B200_X2_ONEWAVE=0" "$a"
done
} > gpurun_out/r02j_ab.txt 2>&1
cat gpurun_out/r02j_ab.txt
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02j_pytest.txt
tail -5 gpurun_out/r02j_pytest.txt
