#!/bin/bash
one() { python bench.py "$@" --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  step-frac %.3f  G/s %.2f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['step']['frac'], d['value']/1e9))"; }
V=$PWD/scratch/variants/libswe_b200_div.so
echo -n "config 2 (1024^2 f64) minmax: "; one --config 2
echo -n "config 2 (1024^2 f64) minmax+div64: "; SWE_B200_LIB=$V one --config 2
echo -n "config 2 at 4096^2 f64 minmax: "; one --config 2 --n 4096
echo -n "config 2 at 4096^2 f64 minmax+div64: "; SWE_B200_LIB=$V one --config 2 --n 4096
echo "--- full gpu suite, default lib"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- fast-policy tests, div64 lib"; SWE_B200_LIB=$V python -m pytest tests -m gpu -x -q -k "fast or golden or large_grid or full_size or lake_at_rest or smoke" 2>&1 | tail -3
