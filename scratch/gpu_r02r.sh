#!/bin/bash
# round 2: parity tables with the final build (default + X2_ACCURATE), accurate-build cost
mkdir -p gpurun_out
V=$PWD/scratch/variants
timeout 900 python scratch/parity_report.py --out gpurun_out/r02r_parity.md > /dev/null 2> gpurun_out/r02r_parity.err
SWE_B200_LIB=$V/libswe_b200_acc.so timeout 600 python scratch/parity_report.py --sizes 5:16384 --out gpurun_out/r02r_parity_acc.md > /dev/null 2> gpurun_out/r02r_parity_acc.err
SWE_B200_LIB=$V/libswe_b200_acc.so python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02r_bench_acc.json 2> gpurun_out/r02r_bench_acc.err
python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench_acc.json').read().strip().splitlines()[-1]); print('acc build ms/step %.3f' % d['ms_per_step'])"
tail -2 gpurun_out/r02r_parity.err gpurun_out/r02r_parity_acc.err
