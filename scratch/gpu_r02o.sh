#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multistrip.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r02o_pytest.txt
tail -4 gpurun_out/r02o_pytest.txt
