#!/bin/bash
# Float64 fast path: side measurements + the full GPU suite
mkdir -p gpurun_out
one() { python bench.py "$@" --steps 20 --warmup 3 --no-cpu 2>/dev/null | tee -a gpurun_out/f64_lines.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f  G/s %.2f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['value']/1e9))"; }
rm -f gpurun_out/f64_lines.json
echo -n "config 2 (1024^2 f64): "; one --config 2
echo -n "config 2 at 4096^2 f64: "; one --config 2 --n 4096
echo "--- full gpu suite"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
