import sys, math
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import cases
from oracle import oracle
P = cases.load_package()
from test_gpu_parity import SMALL_CASES
np.set_printoptions(linewidth=200, precision=6)

def run(name, dtype, strict, fused, nsteps):
    case, state = SMALL_CASES[name](dtype)
    s = P.simulator_from_case(case, strict=strict, fused=fused)
    o = oracle.OracleSim(case)
    s.set_state_interior(*state); o.set_state_interior(*state)
    mts = case.get("maximum_timestep", math.inf)
    first_bad = None
    for n in range(nsteps):
        s.step(mts); o.step(mts)
        if first_bad is None:
            a = s.get_state_full()[:, 2:-2, 2:-2]; b = o.get_state_full()[:, 2:-2, 2:-2]
            e = [cases.rel_linf(a[k], b[k]) for k in range(3)]
            if max(e) > (0 if strict else 1e-6):
                first_bad = (n, e, s.dt, o.dt)
    a = s.get_state_full()[:, 2:-2, 2:-2]; b = o.get_state_full()[:, 2:-2, 2:-2]
    e = [cases.rel_linf(a[k], b[k]) for k in range(3)]
    k = int(np.argmax(e))
    idx = np.unravel_index(np.argmax(np.abs(a[k].astype(float) - b[k])), a[k].shape)
    print(f"{name:18s} {dtype} strict={int(strict)} fused={int(fused)} n={nsteps}: relLinf={['%.2e'%x for x in e]} dt {s.dt:.6g}/{o.dt:.6g} worst field {k} at {idx} "
          f"gpu={a[k][idx]:.8g} orc={b[k][idx]:.8g} first_bad={first_bad}")


for name in sorted(SMALL_CASES):
    for dtype in ("float64", "float32"):
        run(name, dtype, False, True, 100)

# fixed-time comparison
def run_T(name, dtype, nref):
    case, state = SMALL_CASES[name](dtype)
    o = oracle.OracleSim(case); o.set_state_interior(*state)
    mts = case.get("maximum_timestep", math.inf)
    for n in range(nref): o.step(mts)
    T = o.t
    o2 = oracle.OracleSim(case); o2.set_state_interior(*state); o2.simulate_to_time(T, maximum_timestep=case.get("maximum_timestep"))
    s = P.simulator_from_case(case); s.set_state_interior(*state)
    ns = P.simulate_to_time(s, T, maximum_timestep=case.get("maximum_timestep"))
    a = s.get_state_full()[:, 2:-2, 2:-2]; b = o2.get_state_full()[:, 2:-2, 2:-2]
    e = [cases.rel_linf(a[k], b[k]) for k in range(3)]
    extra = ""
    if case.get("B") is not None and not np.isscalar(case.get("B")):
        Bc = cases.synthetic().cell_average(np.asarray(case["B"]))[2:-2, 2:-2]
        ws = (a[0].astype(float) - Bc) >= case["depth_cutoff"]; wo = (b[0].astype(float) - Bc) >= case["depth_cutoff"]
        extra = f" mask mismatch {np.sum(ws != wo)} of {ws.size} wetfrac {wo.mean():.2f}"
    print(f"T-compare {name:18s} {dtype} T={T:.5g} steps gpu {ns}: relLinf={['%.2e'%x for x in e]}" + extra)
for name in sorted(SMALL_CASES):
    for dtype in ("float64", "float32"):
        run_T(name, dtype, 100)
