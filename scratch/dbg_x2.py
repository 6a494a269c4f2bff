"""Compare the x2 (packed) and the generic fast fused kernels against the CPU oracle step by step."""
import os, sys, math
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import cases
from test_gpu_parity import SMALL_CASES
P = cases.load_package(); oracle = cases.load_oracle()
name = sys.argv[1] if len(sys.argv) > 1 else "terrain_runoff"
case, state = SMALL_CASES[name]("float32")
def mk(nox2):
    os.environ["SWE_B200_NO_X2"] = "1" if nox2 else "0"
    s = P.simulator_from_case(case, strict=False, fused=True); s.set_state_interior(*state); return s
a, b = mk(False), mk(True)
o = oracle.OracleSim(case); o.set_state_interior(*state)
mts = case.get("maximum_timestep", math.inf)
I = lambda f: f[:, 2:-2, 2:-2]
for n in range(1, 101):
    a.step(mts); b.step(mts); o.step(mts)
    if n in (1, 2, 3, 5, 10, 25, 50, 100):
        A, B, O = I(a.get_state_full()), I(b.get_state_full()), I(o.get_state_full())
        ea = cases.field_errors(A, O); eb = cases.field_errors(B, O)
        print(n, "x2:", ["%.1e" % e for e in ea], "dt %.3e" % abs(a.dt / o.dt - 1), " generic:", ["%.1e" % e for e in eb], "dt %.3e" % abs(b.dt / o.dt - 1))
        if n == 1:
            for k in range(3):
                d = np.abs(A[k].astype(np.float64) - O[k]); idx = np.unravel_index(np.argmax(d), d.shape)
                print("   worst cell var", k, idx, "x2", A[k][idx], "oracle", O[k][idx], "generic", B[k][idx])
# error map after a single step from the initial state
a2, o2 = mk(False), oracle.OracleSim(case)
o2.set_state_interior(*state); a2.step(mts); o2.step(mts)
A, O = a2.get_state_full(), o2.get_state_full()
d = np.max(np.abs(A.astype(np.float64) - O), axis=0)
bad = d > 1e-5 * max(1.0, np.max(np.abs(O)))
print("full array shape", A.shape, "bad cells", int(bad.sum()))
rows = np.where(bad.any(axis=1))[0]; cols = np.where(bad.any(axis=0))[0]
print("bad rows (axis1 idx):", rows[:40], "...", len(rows)); print("bad cols (axis2 idx):", cols[:40], "...", len(cols))
