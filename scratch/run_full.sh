#!/bin/bash
# full GPU validation of the tree: pytest -m gpu, smoke, bench (N=1), ncu launch list + one --set full capture
mkdir -p gpurun_out
TAG=${TAG:-r01g}
echo "=== pytest -m gpu"; SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "pytest seconds: $SECONDS" | tee -a gpurun_out/${TAG}_pytest.txt
echo "=== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.txt
echo "=== bench"; python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/${TAG}_bench_ref.json
echo "=== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "=== ncu full (K1, K2)"
ncu --set full --clock-control none --import-source on -k regex:k_stage_x2 -s 6 -c 2 -o gpurun_out/${TAG}_x2 -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
