#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02z_pytest.txt
tail -4 gpurun_out/r02z_pytest.txt
timeout 900 python scratch/parity_report.py --out gpurun_out/r02z_parity.md > /dev/null 2> gpurun_out/r02z_parity.err
grep worst gpurun_out/r02z_parity.md
python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/r02z_bench_40steps.json 2>> gpurun_out/r02z_bench.err
python bench.py --config 3 --no-cpu > gpurun_out/r02z_bench_config3.json 2>> gpurun_out/r02z_bench.err
python bench.py --config 4 --no-cpu > gpurun_out/r02z_bench_config4.json 2>> gpurun_out/r02z_bench.err
for f in r02z_bench r02z_bench_40steps r02z_bench_config3 r02z_bench_config4; do python - $f <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[1], 'ms/step %.4f %.2f G/s K1 %.4f K2 %.4f K2-frac %.3f step-frac %.3f e2e %.2f clocks %s %s' % (d['ms_per_step'], d['value']/1e9, r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'], d['clocks']['reasons']))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02z_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 6 -c 2 -f -o gpurun_out/r02z_x2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02z_ncu_x2.log 2>&1
