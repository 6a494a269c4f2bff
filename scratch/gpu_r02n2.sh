#!/bin/bash
# final confirmation on an 8-GPU box: strips on real devices (bitwise tests), strong scaling N = 1, 2, 4, 8
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multistrip.py tests/test_gpu_parity_r2.py -k "strip or close_peers" -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02n2_pytest.txt
tail -2 gpurun_out/r02n2_pytest.txt
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02n2_n1.json 2> gpurun_out/r02n2_n1.err
tr 2 29521 > gpurun_out/r02n2_n2.json 2> gpurun_out/r02n2_n2.err
tr 4 29541 > gpurun_out/r02n2_n4.json 2> gpurun_out/r02n2_n4.err
tr 8 29581 > gpurun_out/r02n2_n8.json 2> gpurun_out/r02n2_n8.err
for f in r02n2_n1 r02n2_n2 r02n2_n4 r02n2_n8; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'N=%d ms/step %.4f  %.2f G/s e2e %.2f G/s launches %d clocks %s %s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
