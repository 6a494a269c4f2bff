#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02w_pytest.txt
tail -4 gpurun_out/r02w_pytest.txt
python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/r02w_bench_config2.json 2> gpurun_out/r02w.err
python bench.py --config 2 --size 4096 --no-cpu > gpurun_out/r02w_bench_config2_4096.json 2>> gpurun_out/r02w.err
python bench.py --dtype float64 --size 8192 --no-cpu > gpurun_out/r02w_bench_config5_f64_8192.json 2>> gpurun_out/r02w.err
python bench.py --config 3 --dtype float64 --no-cpu > gpurun_out/r02w_bench_config3_f64.json 2>> gpurun_out/r02w.err
for f in r02w_bench_config2 r02w_bench_config2_4096 r02w_bench_config5_f64_8192 r02w_bench_config3_f64; do python - $f <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[1], r['kernel_family'], 'ms/step %.4f %.2f G/s K2-frac %.3f step-frac %.3f' % (d['ms_per_step'], d['value']/1e9, r['frac'], r['step']['frac']))
PY
done
ncu --set full --clock-control none --import-source on -k regex:k_stage_d1 -s 6 -c 2 -f -o gpurun_out/r02w_d1 python bench.py --config 2 --size 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02w_ncu_d1.log 2>&1
