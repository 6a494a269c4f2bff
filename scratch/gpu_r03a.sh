#!/bin/bash
# dry-stretch skip (X2_DRYSKIP): exactness test, A/B bench (skip on / off; 54 % wet and fully wet), then the whole GPU suite
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k dry_stretch 2>&1 | tail -30 ) > gpurun_out/r03a_dryskip_pytest.txt
tail -8 gpurun_out/r03a_dryskip_pytest.txt
show() { python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/%s.json' % n).read().strip().splitlines()[-1]); r = d['roofline']
    print('%s ms/step %.4f %.2f G/s K2-frac %.3f step-frac %.3f e2e %.2f G/s launches %d clocks %s %s' % (n, d['ms_per_step'], d['value']/1e9, r['frac'], r['step']['frac'], d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e:
    print(n, 'FAILED', e)
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err; show r03a_bench
SWE_B200_NO_DRYSKIP=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r03a_bench_noskip.json 2> gpurun_out/r03a_bench_noskip.err; show r03a_bench_noskip
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --level 60 > gpurun_out/r03a_bench_wet.json 2> gpurun_out/r03a_bench_wet.err; show r03a_bench_wet
SWE_B200_NO_DRYSKIP=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --level 60 > gpurun_out/r03a_bench_wet_noskip.json 2> gpurun_out/r03a_bench_wet_noskip.err; show r03a_bench_wet_noskip
( timeout 600 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r03a_pytest.txt
tail -4 gpurun_out/r03a_pytest.txt
