#!/bin/bash
# round 2: verification of the final kernels (full GPU suite), bench lines, ncu captures
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 ) > gpurun_out/r02i_pytest.txt
tail -6 gpurun_out/r02i_pytest.txt
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu --steps 20 --warmup 5 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%s ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f launches %d clocks %s %s' % (r['kernel_family'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run "A=1" ""
run "A=1" ""
run "A=1" "--config 2"
run "A=1" "--config 2 --n 4096"
run "A=1" "--config 5 --dtype float64 --n 8192"
run "A=1" "--config 3 --dtype float64"
run "A=1" "--config 3"
run "A=1" "--config 4"
} > gpurun_out/r02i_ab.txt 2>&1
cat gpurun_out/r02i_ab.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err
python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_config2.json 2>> gpurun_out/r02i_bench.err
python bench.py --config 2 --n 4096 --no-cpu > gpurun_out/r02i_bench_config2_4096.json 2>> gpurun_out/r02i_bench.err
python bench.py --config 3 --no-cpu > gpurun_out/r02i_bench_config3.json 2>> gpurun_out/r02i_bench.err
python bench.py --config 4 --no-cpu > gpurun_out/r02i_bench_config4.json 2>> gpurun_out/r02i_bench.err
python bench.py --config 3 --bc neumann --no-cpu > gpurun_out/r02i_bench_config3_neumann.json 2>> gpurun_out/r02i_bench.err
python bench.py --config 4 --bc neumann --no-cpu > gpurun_out/r02i_bench_config4_neumann.json 2>> gpurun_out/r02i_bench.err
python bench.py --level 12 --steps 40 --no-cpu > gpurun_out/r02i_bench_fullywet_40steps.json 2>> gpurun_out/r02i_bench.err
python bench.py --steps 40 --no-cpu > gpurun_out/r02i_bench_40steps.json 2>> gpurun_out/r02i_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02i_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 6 -c 2 -f -o gpurun_out/r02i_x2 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02i_ncu_x2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_d1 -s 6 -c 2 -f -o gpurun_out/r02i_d1 python bench.py --config 2 --n 4096 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02i_ncu_d1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.txt 2>&1; tail -3 gpurun_out/r02i_smoke.txt
