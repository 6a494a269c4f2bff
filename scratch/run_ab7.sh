#!/bin/bash
mkdir -p gpurun_out
for ny in 2048 4096; do
export BENCH_ARGS="--ny $ny"
echo "ny=$ny"
scratch/ab.sh "SWE_B200_X2_WAVES=1" "SWE_B200_X2_WAVES=0" "SWE_B200_X2_WAVES=1 SWE_B200_X2_ROWS=63" "SWE_B200_X2_WAVES=1 SWE_B200_X2_ROWS=90" "SWE_B200_X2_WAVES=1"
done 2>&1 | tee gpurun_out/ab7.txt
export BENCH_ARGS=""
scratch/ab.sh "SWE_B200_X2_WAVES=1" "SWE_B200_X2_WAVES=0" 2>&1 | tee -a gpurun_out/ab7.txt
