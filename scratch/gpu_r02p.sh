#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multistrip.py tests/test_gpu_parity_r2.py -m gpu -q --maxfail=20 2>&1 | tail -30 ) > gpurun_out/r02p_pytest.txt
tail -12 gpurun_out/r02p_pytest.txt
