#!/bin/bash
# 2-GPU check of the warp-parallel CFL all-gather (lane r polls rank r) and the two-thread U-stamp wait in k_tail:
# strip tests (bitwise), N = 2 bench, per-launch trace
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multistrip.py tests/test_gpu_parity_r2.py -k "strip or close_peers" -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02p2_pytest.txt
tail -2 gpurun_out/r02p2_pytest.txt
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02p2_n1.json 2> gpurun_out/r02p2_n1.err
tr 2 29521 > gpurun_out/r02p2_n2.json 2> gpurun_out/r02p2_n2.err
SWE_B200_TRACE=1 tr 2 29523 --rebalance 0 > gpurun_out/r02p2_n2_trace.json 2> gpurun_out/r02p2_n2_trace.err
grep "swe_b200 trace" gpurun_out/r02p2_n2_trace.err | tail -4
for f in r02p2_n1 r02p2_n2 r02p2_n2_trace; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'N=%d ms/step %.4f  %.2f G/s e2e %.2f G/s launches %d clocks %s %s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
