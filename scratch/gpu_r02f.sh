#!/bin/bash
# round 2, single-GPU call: full GPU suite, FILL16 / d1 occupancy A/B, parity tables, ncu captures (x2 + d1), side configs
mkdir -p gpurun_out
V=$PWD/scratch/variants
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 ) > gpurun_out/r02f_pytest.txt
tail -6 gpurun_out/r02f_pytest.txt
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f clocks %s %s' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run "A=1" ""
run "SWE_B200_LIB=$V/libswe_b200_fill16.so" ""
run "A=1" ""
run "SWE_B200_LIB=$V/libswe_b200_fill16.so" ""
run "A=1" "--steps 40"
run "SWE_B200_LIB=$V/libswe_b200_fill16.so" "--steps 40"
run "A=1" "--config 2"
run "SWE_B200_LIB=$V/libswe_b200_d1m3.so" "--config 2"
run "A=1" "--config 2 --n 4096"
run "SWE_B200_LIB=$V/libswe_b200_d1m3.so" "--config 2 --n 4096"
run "A=1" "--config 3"
run "A=1" "--config 4"
run "A=1" "--level 12 --steps 40"
} > gpurun_out/r02f_ab.txt 2>&1
cat gpurun_out/r02f_ab.txt
timeout 900 python scratch/parity_report.py --out gpurun_out/r02f_parity.md > /dev/null 2> gpurun_out/r02f_parity.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 6 -c 2 -f -o gpurun_out/r02f_x2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02f_ncu_x2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_d1 -s 6 -c 2 -f -o gpurun_out/r02f_d1 python bench.py --config 2 --n 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02f_ncu_d1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
