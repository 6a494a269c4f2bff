#!/bin/bash
# A/B timing of the stage kernels under env knobs: prints ms/step, K1 ms, K2 ms per configuration
# usage: [BENCH_ARGS="--ny 2048"] ab.sh "ENV=..." ...
run() {
  echo -n "$1: "
  env $1 python bench.py --no-cpu --steps 10 --warmup 3 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.3f  K1 %.3f  K2 %.3f  step-frac %.3f  G/s %.2f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['value']/1e9))"
}
for cfg in "$@"; do run "$cfg"; done
