#!/bin/bash
# round 2, final 8-GPU call: strong scaling N = 1, 2, 4, 8 on one box with the final build
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02n_n1.json 2> gpurun_out/r02n_n1.err
tr 2 29521 > gpurun_out/r02n_n2.json 2> gpurun_out/r02n_n2.err
tr 4 29541 > gpurun_out/r02n_n4.json 2> gpurun_out/r02n_n4.err
tr 8 29581 > gpurun_out/r02n_n8.json 2> gpurun_out/r02n_n8.err
for f in r02n_n1 r02n_n2 r02n_n4 r02n_n8; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'N=%d ms/step %.4f  %.2f G/s e2e %.2f G/s launches %d clocks %s %s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']), [(p['k1_ms'],p['k2_ms']) for p in (d.get('per_rank') or [])][:8])
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
