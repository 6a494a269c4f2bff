#!/bin/bash
mkdir -p gpurun_out
one() { python bench.py "$@" --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f  G/s %.2f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['value']/1e9))"; }
for r in 64 32 16 8 auto; do echo -n "config 2 rows=$r: "; if [ $r = auto ]; then one --config 2; else SWE_B200_FUSED_ROWS=$r one --config 2; fi; done
for r in 126 auto 30; do echo -n "config 3 big=$r: "; if [ $r = auto ]; then one --config 3; else SWE_B200_X2_ROWS=$r one --config 3; fi; done
echo -n "config 4: "; one --config 4
echo -n "config 5 n=8192: "; one --n 8192
echo -n "config 2 n=4096 f64: "; one --config 2 --n 4096
