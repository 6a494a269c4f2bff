#!/bin/bash
# first GPU call of this session: A/B of the x2 variants at 8192^2, parity of the candidates
mkdir -p gpurun_out
export BENCH_ARGS="--n 8192"
V=$PWD/scratch/variants
scratch/ab.sh "SWE_B200_LIB=$V/libswe_b200_old.so" "SWE_B200_LIB=$V/libswe_b200_base.so" "SWE_B200_LIB=$V/libswe_b200_stcs.so" "SWE_B200_LIB=$V/libswe_b200_tcarry.so" \
  "SWE_B200_LIB=$V/libswe_b200_fric3.so" "SWE_B200_LIB=$V/libswe_b200_all.so" "SWE_B200_LIB=$V/libswe_b200_base.so" 2>&1 | tee gpurun_out/ab1.txt
echo "--- parity: default lib (golden + fast parity)"
timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -q -m gpu -x -k "golden or fast_fused_parity or lake_at_rest or large_grid" 2>&1 | tail -5 | tee gpurun_out/par_default.txt
echo "--- parity: all-variant lib"
SWE_B200_LIB=$V/libswe_b200_all.so timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -q -m gpu -x -k "golden or fast_fused_parity or lake_at_rest or large_grid" 2>&1 | tail -5 | tee gpurun_out/par_all.txt
