#!/bin/bash
mkdir -p gpurun_out
export BENCH_ARGS="--n 8192"
V=$PWD/scratch/variants
args=()
for v in d m3 m3n m3nt d; do args+=("SWE_B200_LIB=$V/libswe_b200_$v.so"); done
scratch/ab.sh "${args[@]}" 2>&1 | tee gpurun_out/ab5.txt
