#!/bin/bash
mkdir -p gpurun_out
export BENCH_ARGS="--n 8192"
V=$PWD/scratch/variants
args=()
for v in d d_cu d_cu_t3 d; do args+=("SWE_B200_LIB=$V/libswe_b200_$v.so"); done
scratch/ab.sh "${args[@]}" 2>&1 | tee gpurun_out/ab3.txt
export BENCH_ARGS=""
scratch/ab.sh "SWE_B200_LIB=$V/libswe_b200_d.so" "SWE_B200_LIB=$V/libswe_b200_d_cu.so" 2>&1 | tee -a gpurun_out/ab3.txt
