#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -30 ) > gpurun_out/r02k_pytest.txt
tail -5 gpurun_out/r02k_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
