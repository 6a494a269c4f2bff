#!/bin/bash
mkdir -p gpurun_out
export BENCH_ARGS="--n 8192"
V=$PWD/scratch/variants
args=()
for v in o o_t2 o_f o_t2f o_t2fs n_t2f n_t2fs o; do args+=("SWE_B200_LIB=$V/libswe_b200_$v.so"); done
scratch/ab.sh "${args[@]}" 2>&1 | tee gpurun_out/ab2.txt
