#!/bin/bash
# X2_NOGHOST_FAST=2 (both FAST flavours in K2): bitwise check against the shipped build, then a short A/B
mkdir -p gpurun_out
V=$PWD/scratch/variants
C=lake_wall,lake_periodic_films,lake_neumann,urban_norecon
SWE_B200_LIB=$V/libswe_b200_ng2.so timeout 60 python scratch/dump_state.py gpurun_out/r03d_state_ng2.npz $C 2> gpurun_out/r03d_dump_ng2.err &
timeout 60 python scratch/dump_state.py gpurun_out/r03d_state_ship.npz $C 2> gpurun_out/r03d_dump_ship.err &
wait
python scratch/dump_state.py --compare gpurun_out/r03d_state_ship.npz gpurun_out/r03d_state_ng2.npz > gpurun_out/r03d_bitwise.txt 2>&1; cat gpurun_out/r03d_bitwise.txt
rm -f gpurun_out/r03d_state_*.npz
run() { echo -n "$1 $3: "; env $2 timeout 60 python bench.py --no-cpu --steps 20 --warmup 5 $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f clocks %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"; }
{
run ng2 "SWE_B200_LIB=$V/libswe_b200_ng2.so" ""
run ship "A=1" ""
run ng2 "SWE_B200_LIB=$V/libswe_b200_ng2.so" "--config 3"
run ship "A=1" "--config 3"
} > gpurun_out/r03d_ab.txt 2>&1
cat gpurun_out/r03d_ab.txt
