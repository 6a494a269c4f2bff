#!/bin/bash
# round 2, first GPU call: regression tests, parity tables (default + accurate build), sustained / fully-wet bench lines
mkdir -p gpurun_out
V=$PWD/scratch/variants
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r02a_pytest.txt
timeout 900 python scratch/parity_report.py --out gpurun_out/r02a_parity.md > /dev/null 2> gpurun_out/r02a_parity.err
SWE_B200_LIB=$V/libswe_b200_acc.so timeout 600 python scratch/parity_report.py --sizes 5:16384 --out gpurun_out/r02a_parity_acc.md > /dev/null 2> gpurun_out/r02a_parity_acc.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/r02a_bench_40.json 2>> gpurun_out/r02a_bench.err
python bench.py --steps 40 --warmup 5 --no-cpu --level 12 > gpurun_out/r02a_bench_wet40.json 2>> gpurun_out/r02a_bench.err
SWE_B200_LIB=$V/libswe_b200_acc.so python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02a_bench_acc.json 2>> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_pytest.txt
for f in r02a_bench r02a_bench_40 r02a_bench_wet40 r02a_bench_acc; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms/step %.3f K1 %.3f K2 %.3f step-frac %.3f e2e %.2f G/s clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
