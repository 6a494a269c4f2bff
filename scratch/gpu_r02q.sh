#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02q_pytest.txt
tail -4 gpurun_out/r02q_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f %.2f G/s K2-frac %.3f step-frac %.3f e2e %.2f G/s launches %d clocks %s %s cpu %.2f M/s' % (d['ms_per_step'], d['value']/1e9, r['frac'], r['step']['frac'], d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['cpu_baseline']['value']/1e6))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
