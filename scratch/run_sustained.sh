#!/bin/bash
# sustained (power-capped) throughput of kernel variants: 30 warm-up + 40 timed steps
V=$PWD/scratch/variants
one() { python bench.py --steps 40 --warmup 30 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; c=d['clocks']
print('ms/step %.3f  K1 %.3f  K2 %.3f  sm_mhz %s min %s power %s %s' % (d['ms_per_step'], r['k1']['avg_launch_ms'], r['avg_launch_ms'], c['sm_mhz'], c['sm_min_mhz'], c['power_w_max'], c['reasons']))"; }
for v in cur nog stcs m4 cur; do echo -n "$v: "; SWE_B200_LIB=$V/libswe_b200_$v.so one; done
