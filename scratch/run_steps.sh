#!/bin/bash
# does the per-step time depend on how many steps are timed / how far the simulation has evolved?
one() { python bench.py "$@" --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.3f  K1 %.3f  K2 %.3f  clocks %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], d['clocks']))"; }
echo -n "steps 10 warm 3: "; one --steps 10 --warmup 3
echo -n "steps 20 warm 3: "; one --steps 20 --warmup 3
echo -n "steps 40 warm 3: "; one --steps 40 --warmup 3
echo -n "steps 10 warm 30: "; one --steps 10 --warmup 30
echo -n "steps 10 warm 3: "; one --steps 10 --warmup 3
