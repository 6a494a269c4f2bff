#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multistrip.py tests/test_gpu_parity_r2.py -k "strip or close_peers" -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r02l_pytest.txt
tail -3 gpurun_out/r02l_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02l_n2.json 2> gpurun_out/r02l_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02l_n2_ref.json 2> gpurun_out/r02l_n2_ref.err
python - <<'PY'
import json
for f in ("r02l_n2","r02l_n2_ref"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get('impl','ours'), 'N=%d ms/step %.4f %.2f G/s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9), d['config'].get('strip_balance'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
tail -3 gpurun_out/r02l_n2.err
