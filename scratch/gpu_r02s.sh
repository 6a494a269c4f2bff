#!/bin/bash
# round 2: burst vs sustained sweep of the bench line; Float64 strips on 2 GPUs
mkdir -p gpurun_out
run() { echo -n "$1 :: $2 : "; env $1 python bench.py --no-cpu $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; c=d['clocks']
print('%s N=%d ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f e2e %.2f G/s sm_mhz median %s min %s power %s %s' % (r['kernel_family'], d['n_gpus'], d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['e2e']['value']/1e9, c['sm_mhz'], c.get('sm_min_mhz'), c.get('power_w_max'), c['reasons']))"; }
{
for s in 10 20 40 100; do run "A=1" "--steps $s --warmup 5"; done
run "A=1" "--steps 100 --warmup 10 --level 12"
run "A=1" "--config 5 --dtype float64 --n 8192 --steps 20 --warmup 5"
} > gpurun_out/r02s_sweep.txt 2>&1
cat gpurun_out/r02s_sweep.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --dtype float64 --n 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02s_f64_n2.json 2> gpurun_out/r02s_f64_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02s_f64_n2.json').read().strip().splitlines()[-1]); r=d['roofline']
print('f64 8192^2 N=2:', r['kernel_family'], 'ms/step %.4f %.2f G/s' % (d['ms_per_step'], d['value']/1e9), d.get('per_rank'))" || tail -5 gpurun_out/r02s_f64_n2.err
