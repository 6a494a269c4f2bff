#!/bin/bash
mkdir -p gpurun_out
python bench.py --dtype float64 --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02u_f64_n1.json 2> gpurun_out/r02u_f64_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --dtype float64 --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02u_f64_n2.json 2> gpurun_out/r02u_f64_n2.err
for f in r02u_f64_n1 r02u_f64_n2; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], r['kernel_family'], 'N=%d ms/step %.4f %.2f G/s step-frac %.3f' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, r['step']['frac']), d.get('per_rank'))
except Exception as e: print(sys.argv[1], 'FAILED', e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
