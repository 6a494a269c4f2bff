#!/usr/bin/env python
"""Parity report of the Float32 performance path (x2 kernels), written as markdown (profiles/r02_parity*.md).

Part A  small instances of every BASELINE config: fast CUDA path vs the CPU oracle after 20 / 25 / 100 steps,
        beside the oracle's own sensitivity to a +-1 ulp perturbation of the initial state (4 seeds) and the
        strict CUDA path (bit-exact).
Part B  FULL-SIZE configs: fast vs strict CUDA path per cell, compared on the device (swe_b200_compare_states),
        beside strict(+-1 ulp perturbed input) vs strict(original) = the conditioning of the case at that size.

  python scratch/parity_report.py [--out FILE] [--small-only] [--full-only] [--sizes 16384,4096,8192,1024]
The library under test is the one SWE_B200_LIB names (default: the in-tree build), so the same script produces the
table of the X2_ACCURATE build.
"""
import argparse
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402

P = cases.load_package()


def interior(full, gc=2):
    return full[:, gc:-gc, gc:-gc]


def wet_mask(case, w):
    B = case.get("B")
    if B is None or np.isscalar(B):
        return w.astype(np.float64) >= case.get("depth_cutoff", 1e-5)
    Bc = cases.synthetic().cell_average(np.asarray(B))[2:-2, 2:-2]
    return (w.astype(np.float64) - Bc) >= case["depth_cutoff"]


def fmt(v):
    return " / ".join("%.1e" % x for x in v)


def small(out, checkpoints=(20, 25, 100), seeds=(1, 2, 3, 4)):
    import test_gpu_parity as T
    oracle = cases.load_oracle()
    out.write("## A. Small instances: fast Float32 CUDA path vs the CPU oracle\n\n"
              "rel-L-inf per field (w / hu / hv; w against max|w|, momenta against max(|hu|,|hv|)), interior cells.\n"
              "`sens` = the ORACLE run from a +-1 ulp perturbed initial state against the oracle itself (max over "
              f"{len(seeds)} seeds): what the scheme does to the smallest input change Float32 can express.\n"
              "`mask` = cells whose wet flag differs (CUDA fast vs oracle) / (perturbed oracle vs oracle, max over seeds).\n"
              "`err / max(1e-5, sens)`: w against its own sensitivity, both momenta against the larger of the two momentum sensitivities "
              "(they share one scale; per field the sensitivity of four seeds is heavy-tailed).\n\n"
              "| case | steps | fast vs oracle | oracle 1-ulp sens | err / max(1e-5, sens) | mask | dt rel err | strict vs oracle |\n"
              "|---|---|---|---|---|---|---|---|\n")
    worst = 0.0
    for name in sorted(T.SMALL_CASES):
        if name == "lake_step_fe":
            continue
        case, state = T.SMALL_CASES[name]("float32")
        mts = case.get("maximum_timestep", math.inf)
        ref, dts_o = T.run_oracle(oracle, case, state, checkpoints)
        pert = [T.run_oracle(oracle, case, cases.perturb_ulp(state, "float32", seed=sd), checkpoints) for sd in seeds]
        sims = {k: P.simulator_from_case(case, strict=(k == "strict"), fused=True) for k in ("fast", "strict")}
        for s in sims.values():
            s.set_state_interior(*state)
        dts = []
        for n in range(1, max(checkpoints) + 1):
            for s in sims.values():
                s.step(mts)
            dts.append(sims["fast"].dt)
            if n not in checkpoints:
                continue
            a, b = interior(sims["fast"].get_state_full()), ref[n]
            st = interior(sims["strict"].get_state_full())
            err = cases.field_errors(a, b)
            sens = np.max([cases.field_errors(pp[n], b) for pp, _ in pert], axis=0)
            sens_m = max(sens[1], sens[2])          # the two momenta share one scale and one yardstick (heavy-tailed per field)
            ratio = max(e / max(1e-5, sn) for e, sn in zip(err, (sens[0], sens_m, sens_m)))
            worst = max(worst, ratio)
            mism = int(np.sum(wet_mask(case, a[0]) != wet_mask(case, b[0])))
            mism_ref = max(int(np.sum(wet_mask(case, pp[n][0]) != wet_mask(case, b[0]))) for pp, _ in pert)
            err_dt = float(np.max(np.abs(np.array(dts) - dts_o[:n]) / dts_o[:n]))
            strict_err = "bit-exact" if np.array_equal(st, b) else fmt(cases.field_errors(st, b))
            out.write(f"| {name} ({sims['fast'].kernel_family()}) | {n} | {fmt(err)} | {fmt(sens)} | {ratio:.2f} | {mism} / {mism_ref} | "
                      f"{err_dt:.1e} | {strict_err} |\n")
            out.flush()
    out.write(f"\nworst err / max(1e-5, sens) over the table: **{worst:.2f}**\n\n")


FULL = {
    5: ("configs[4] lake, bh2021, minmod", lambda n: cases.lake_case(n=n, dtype="float32", level=0.0), (10, 20)),
    3: ("configs[2] terrain, rain + Horton, bsa2012", lambda n: cases.terrain_case(n=n, dtype="float32", level=1.0), (10, 20, 100)),
    4: ("configs[3] urban, NoReconstruction, bsa2012", lambda n: cases.urban_case(n=n, dtype="float32", level=1.0), (10, 20, 100)),
    2: ("configs[1] SWASHES 4.1.2 dam break (Float64)", lambda n: cases.swashes_dambreak_2d(n=n, dtype="float64"), (20, 100, 200)),
}


def full(out, sizes):
    out.write("## B. Full-size configs: fast vs strict CUDA path, per cell, compared on the device\n\n"
              "The strict path is bit-exact against the CPU oracle wherever the oracle can run (tests/test_gpu_parity.py), so "
              "at these sizes it stands in for the reference.  `sens` = strict run from a +-1 ulp perturbed initial state "
              "against the strict run from the original one.\n\n"
              "`cells > tol` = number of cells whose error exceeds the north_star tolerance x the field's scale (fast / perturbed strict).\n\n"
              "| config | grid | steps | fast vs strict (w / hu / hv) | strict 1-ulp sens | err / max(tol, sens) | mask mismatches fast / perturbed (of wet cells) | cells > tol, fast (w / hu / hv) | cells > tol, perturbed strict | cells > 10 tol, fast | cells > 10 tol, perturbed strict | dt rel diff |\n"
              "|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for cfg, n in sizes:
        label, build, checkpoints = FULL[cfg]
        t0 = time.perf_counter()
        case, state = build(n)
        dtype = case["dtype"]
        tol = 1e-5 if dtype == "float32" else 1e-12
        mts = case.get("maximum_timestep")
        fast = P.simulator_from_case(case, strict=False, fused=True)
        strict = P.simulator_from_case(case, strict=True, fused=True)
        pert = P.simulator_from_case(case, strict=True, fused=True)
        fast.set_state_interior(*state)
        strict.set_state_interior(*state)
        pert.set_state_interior(*cases.perturb_ulp(state, dtype, seed=1))
        del state
        done = 0
        for n_steps in checkpoints:
            for s in (fast, strict, pert):
                s.run_steps(n_steps - done, mts)
            done = n_steps
            c = fast.compare_states(strict, tol=tol)
            cp = pert.compare_states(strict, tol=tol)
            c10 = fast.compare_states(strict, tol=10 * tol)["cells_over_tol"]
            cp10 = pert.compare_states(strict, tol=10 * tol)["cells_over_tol"]
            sm = max(cp["rel_linf"][1], cp["rel_linf"][2])
            ratio = max(e / max(tol, sn) for e, sn in zip(c["rel_linf"], (cp["rel_linf"][0], sm, sm)))
            ddt = abs(fast.dt - strict.dt) / strict.dt
            out.write(f"| {label} ({fast.kernel_family()}) | {n}x{n} {dtype} | {n_steps} | {fmt(c['rel_linf'])} | {fmt(cp['rel_linf'])} | "
                      f"{ratio:.2f} | {c['mask_mismatch']} / {cp['mask_mismatch']} (of {c['wet_other']}) | "
                      f"{' / '.join(map(str, c['cells_over_tol']))} | {' / '.join(map(str, cp['cells_over_tol']))} | "
                      f"{' / '.join(map(str, c10))} | {' / '.join(map(str, cp10))} | {ddt:.1e} |\n")
            out.flush()
            assert c["nan_cells"] == 0
        print(f"config {cfg} at {n}: {time.perf_counter() - t0:.1f} s", file=sys.stderr)
        del fast, strict, pert
    out.write("\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "r02_parity.md"))
    ap.add_argument("--small-only", action="store_true")
    ap.add_argument("--full-only", action="store_true")
    ap.add_argument("--sizes", default="5:16384,3:4096,4:8192,2:1024")
    args = ap.parse_args()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    with open(args.out, "w") as out:
        out.write(f"# Parity of the fast arithmetic policies -- library: `{os.environ.get('SWE_B200_LIB', 'in-tree libswe_b200.so')}`\n\n")
        if not args.full_only:
            small(out)
        if not args.small_only:
            full(out, [tuple(int(x) for x in s.split(":")) for s in args.sizes.split(",")])
    print(open(args.out).read())


if __name__ == "__main__":
    main()
