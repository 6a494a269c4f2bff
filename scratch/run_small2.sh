#!/bin/bash
mkdir -p gpurun_out
one() { python bench.py "$@" --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.4f  K1 %.4f  K2 %.4f  K2-frac %.3f step-frac %.3f  G/s %.2f' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], r['frac'], r['step']['frac'], d['value']/1e9))"; }
for r in auto 18 20 9 12; do echo -n "config 2 rows=$r: "; if [ $r = auto ]; then one --config 2; else SWE_B200_FUSED_ROWS=$r one --config 2; fi; done
echo -n "config 2 n=2048 auto: "; one --config 2 --n 2048
echo -n "config 2 n=2048 rows 64: "; SWE_B200_FUSED_ROWS=64 one --config 2 --n 2048
echo -n "config 2 n=512 auto: "; one --config 2 --n 512
python -m pytest tests -m gpu -x -q -k "strict_steps or golden or swashes" 2>&1 | tail -3
