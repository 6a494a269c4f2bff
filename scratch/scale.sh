#!/bin/bash
# strong-scaling sweep on one box: prints G cell-updates/s and ms/step for each N
for n in "$@"; do
  if [ "$n" = "1" ]; then out=$(python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu 2>/dev/null | tail -1)
  else out=$(python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>/dev/null | tail -1); fi
  echo "$out" > gpurun_out/scale_n$n.json
  echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d  %.2f G/s  %.3f ms/step  step-frac %.3f' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['step']['frac']))"
done
