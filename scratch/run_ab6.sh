#!/bin/bash
mkdir -p gpurun_out
export BENCH_ARGS="--n 8192"
V=$PWD/scratch/variants
args=()
for v in d e e_cu3 e_s e_t3 e; do args+=("SWE_B200_LIB=$V/libswe_b200_$v.so"); done
scratch/ab.sh "${args[@]}" 2>&1 | tee gpurun_out/ab6.txt
scratch/ab.sh "SWE_B200_LIB=$V/libswe_b200_e.so SWE_B200_X2_ROWS=252" "SWE_B200_LIB=$V/libswe_b200_e.so SWE_B200_X2_ROWS=63" 2>&1 | tee -a gpurun_out/ab6.txt
