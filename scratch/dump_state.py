#!/usr/bin/env python
"""A/B of two builds of the library (SWE_B200_LIB=...): runs a few fused Float32 cases and writes the final states + (t, dt)
to an .npz; scratch/dump_state.py --compare a.npz b.npz reports whether they are bitwise equal."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

def cases(P):
    W = P.workloads
    rng = np.random.default_rng(5)
    out = []
    c, s = W.lake_case(n=1536, ny=1100, dtype="float32", level=0.0, wavelength=256.0)
    out.append(("lake_wall", c, s, 80))
    c, s = W.lake_case(n=1000, ny=700, dtype="float32", level=0.0, wavelength=256.0)
    c["bc"] = "periodic"
    w, hu, hv = [np.array(a) for a in s]
    Bc = P.synthetic.cell_average(c["B"])[2:-2, 2:-2].astype(np.float32)
    dry = (w - Bc) < c["depth_cutoff"]
    pick = dry & (rng.uniform(size=w.shape) < 0.2)            # films around the threshold and momenta on dry land
    w = np.where(pick, w + rng.uniform(0, 2 * c["depth_cutoff"], size=w.shape), w).astype(np.float32)
    hu = np.where(pick, 1e-4 * rng.standard_normal(w.shape), 0.0).astype(np.float32)
    out.append(("lake_periodic_films", c, (w, hu, hv), 60))
    c, s = W.lake_case(n=900, ny=640, dtype="float32", level=0.0, wavelength=256.0)
    c["bc"] = "neumann"
    out.append(("lake_neumann", c, s, 60))
    c, s = W.lake_case(n=900, ny=640, dtype="float32", level=0.0, wavelength=256.0)
    c["stepper"] = "fe"; c["friction"] = None
    out.append(("lake_fe_nofric", c, s, 60))
    c, s = W.terrain_case(n=1024, dtype="float32", level=1.0)
    out.append(("terrain_rain_horton", c, s, 40))
    c, s = W.urban_case(n=1024, dtype="float32", level=1.0)
    out.append(("urban_norecon", c, s, 40))
    return out

def main():
    if sys.argv[1] == "--compare":
        a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
        ok = True
        for k in a.files:
            same = np.array_equal(a[k], b[k])
            ok &= same
            d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            print(f"{k:32s} {'bitwise equal' if same else 'DIFFERENT'}  max |diff| {float(np.nanmax(d)):.3e}  cells differing {int(np.sum(a[k] != b[k]))}")
        print("ALL BITWISE EQUAL" if ok else "NOT EQUAL")
        return
    from cases import load_package
    P = load_package()
    res = {}
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None      # optional: a comma-separated subset of the case names
    for name, case, state, steps in cases(P):
        if only and name not in only: continue
        sim = P.simulator_from_case(case, strict=False)
        sim.set_state_interior(*state)
        sim.run_steps(steps)
        res[name + ".state"] = np.stack(sim.get_state_full())
        res[name + ".t_dt"] = np.array([sim.t, sim.dt])
        print(name, sim.kernel_family(), "t", sim.t, file=sys.stderr)
    np.savez(sys.argv[1], **res)

main()
