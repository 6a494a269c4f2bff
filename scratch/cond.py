import sys, math
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import cases
from oracle import oracle
from test_gpu_parity import SMALL_CASES
def run(case, state, n):
    o = oracle.OracleSim(case); o.set_state_interior(*state)
    mts = case.get("maximum_timestep", math.inf)
    for _ in range(n): o.step(mts)
    return o.get_state_full()[:, 2:-2, 2:-2], o.t
for name in sorted(SMALL_CASES):
    for dtype in ("float64", "float32"):
        case, state = SMALL_CASES[name](dtype)
        a, ta = run(case, state, 100)
        b, tb = run(case, cases.perturb_ulp(state, dtype), 100)
        print(f"{name:18s} {dtype}: sensitivity {['%.1e' % e for e in cases.field_errors(b, a)]}  t {ta:.6g} vs {tb:.6g}")
