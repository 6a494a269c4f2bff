#!/bin/bash
# final artefacts of the last build (X2_SUBSAT=1): GPU suite, smoke, bench line (+ 40 steps, + every cell wet), ncu launch list
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r03z_pytest.txt
tail -3 gpurun_out/r03z_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r03z_bench.json 2> gpurun_out/r03z_bench.err
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/r03z_bench_40steps.json 2>> gpurun_out/r03z_bench.err
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu --level 60 > gpurun_out/r03z_bench_fullywet_40steps.json 2>> gpurun_out/r03z_bench.err
for f in r03z_bench r03z_bench_40steps r03z_bench_fullywet_40steps; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms/step %.4f %.2f G/s K1 %.4f K2 %.4f K2-frac %.3f step-frac %.3f e2e %.2f clocks %s %s' % (d['ms_per_step'], d['value']/1e9, r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r03z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03z_ncu_bench.log 2>&1
tail -3 gpurun_out/r03z_launches.csv | cut -c1-200
