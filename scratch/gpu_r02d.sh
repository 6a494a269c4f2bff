#!/bin/bash
# round 2, 2-GPU call: strips on two devices (bitwise tests), strong scaling N = 1, 2 with the merged tail launch,
# the device-side rendezvous and the measured strip rebalancing; per-launch trace at N = 2
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multistrip.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r02d_pytest.txt
tail -3 gpurun_out/r02d_pytest.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02d_n1.json 2> gpurun_out/r02d_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02d_n2.json 2> gpurun_out/r02d_n2.err
SWE_B200_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --rebalance 0 > gpurun_out/r02d_n2_trace.json 2> gpurun_out/r02d_n2_trace.err
grep "swe_b200 trace" gpurun_out/r02d_n2_trace.err | tail -4
for f in r02d_n1 r02d_n2 r02d_n2_trace; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'N=%d ms/step %.4f  %.2f G/s e2e %.2f G/s launches %d' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches']), d.get('per_rank'), d['config'].get('strip_balance'))
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
