#!/bin/bash
# side measurements: BASELINE configs 2-4 through bench.py (same instrumentation as the bench line)
mkdir -p gpurun_out
for c in 2 3 4; do
  python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 3 2>gpurun_out/cfg$c.err | tee gpurun_out/${TAG:-r01i}_bench_config$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('config $c: ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f  G/s %.2f  e2e G/s %.2f  launches %d cpu %.2f M/s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['cpu_baseline']['value']/1e6))"
  tail -2 gpurun_out/cfg$c.err
done
python bench.py --config 2 --steps 20 --warmup 3 --no-cpu --strict 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('config 2 strict: ms/step %.4f  K1 %.4f  K2 %.4f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms']))"
