#!/bin/bash
# 8-GPU box, build with the warp-parallel CFL all-gather: strip tests, strong scaling N = 1, 4, 8 and a per-launch trace at N = 8
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multistrip.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02q2_pytest.txt
tail -2 gpurun_out/r02q2_pytest.txt
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02q2_n1.json 2> gpurun_out/r02q2_n1.err
tr 8 29581 > gpurun_out/r02q2_n8.json 2> gpurun_out/r02q2_n8.err
tr 4 29541 > gpurun_out/r02q2_n4.json 2> gpurun_out/r02q2_n4.err
tr 2 29521 > gpurun_out/r02q2_n2.json 2> gpurun_out/r02q2_n2.err
SWE_B200_TRACE=1 tr 8 29583 > gpurun_out/r02q2_n8_trace.json 2> gpurun_out/r02q2_n8_trace.err
grep "swe_b200 trace" gpurun_out/r02q2_n8_trace.err | tail -8
for f in r02q2_n1 r02q2_n2 r02q2_n4 r02q2_n8 r02q2_n8_trace; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'N=%d ms/step %.4f  %.2f G/s e2e %.2f G/s launches %d clocks %s %s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
