#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/.

  python tests/golden/make_golden.py          # rewrites trajectories.npz + reference_known_answers.json

Two kinds of fixtures:

1. reference_known_answers.json -- the known-answer vectors the REFERENCE'S OWN TESTS hold for this path,
   transcribed as data with the test file:line they come from (paths relative to /root/reference/test).
   The oracle and the CUDA path are both checked against them.
2. trajectories.npz -- outputs of the CPU oracle (oracle/swe_oracle.cpp, the restatement of the reference
   algorithm) on small seeded cases: interior state after N steps and the dt sequence, Float64 and Float32.
   The reference is pure Julia and no Julia toolchain exists in this image, so these are NOT reference-produced
   numbers; they freeze the restatement (any later change of the oracle or of the seeded generators shows up
   as a golden mismatch) and give the GPU tests a committed vector to compare with at run time on the GPU box,
   where /root/reference does not exist.

Closing the gap on a Julia-equipped machine (see tests/golden/dump_reference.jl):
  python tests/golden/make_golden.py --export-inputs DIR ; julia dump_reference.jl DIR ;
  python tests/golden/make_golden.py --import-reference DIR      # -> reference_trajectories.npz, picked up by test_golden.py
"""
from __future__ import annotations

import hashlib
import json
import math
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402

NSTEPS = 25

# small instances of the BASELINE configs / reference test set-ups (tests/cases.py)
GOLDEN_CASES = {
    "swashes412_rk2": lambda dt: cases.swashes_dambreak_2d(n=48, dtype=dt),                       # config 2
    "terrain_shore": lambda dt: cases.terrain_case(n=48, dtype=dt, level=0.0, wavelength=32.0,    # config 3
                                                   friction="bh2021"),
    "urban_norecon": lambda dt: cases.urban_case(n=48, dtype=dt, level=1.0, wavelength=32.0),     # config 4
    "lake_bh2021": lambda dt: cases.lake_case(n=52, ny=44, dtype=dt, level=0.0, wavelength=24.0),  # config 5
    "dry_states_rk2": lambda dt: cases.dry_states_2d(dtype=dt, n=40),            # test_shallow_water_dry_states_2d.jl
    "lake_at_rest_step": lambda dt: cases.lake_at_rest_2d("step", dtype=dt),     # test_lake_at_rest_2d.jl
}

# friction laws with cbrt are not correctly rounded anywhere (glibc / CUDA / Julia): compared at 1e-13, not bitwise
CBRT_CASES = ("urban_norecon",)

KNOWN_ANSWERS = {
    "_doc": "Known-answer vectors held by the reference's own tests (test/*.jl), transcribed as data.",
    "minmod": {
        "source": "test_minmod.jl:24-27",
        "cases": [[1, 2, 3, 1], [-1, -2, -3, -1], [1, -2, 3, 0], [1, 3, 4, 1], [-2, -5, -3, -2], [-3, 2, 4, 0]],
    },
    "bottom_topography_2d": {
        "source": "test_bottom_topography.jl:76-106",
        "setup": "B(corner) = x + y over cell_faces(grid2D, interior=false); see tests/test_golden.py for the grid",
        "B_cell": [[3, 3, 8.0], [3, 4, 18.0], [4, 3, 14.0], [4, 4, 24.0]],
        "B_face_left_x": [[3, 3, 5.0], [4, 4, 21.0]],
        "B_face_right_x": [[3, 3, 11.0], [4, 4, 27.0]],
        "B_face_left_y": [[3, 3, 3.0], [4, 4, 19.0]],
        "B_face_right_y": [[3, 3, 13.0], [4, 4, 29.0]],
        "atol": 1e-10,
    },
    "time_dependent_rain": {
        "source": "test_rain.jl:60-65",
        "rates": [0.015, 0.0003, 0.02], "times": [0.0, 3600.0, 7200.0],
        "queries": [[10.0, 0.015 / 3600.0], [3610.0, 0.0003 / 3600.0], [7210.0, 0.02 / 3600.0],
                    [1.0e6, 0.02 / 3600.0], [0.0, 0.015 / 3600.0]],
    },
    "constant_rain_run": {
        "source": "test_rain.jl:67-77", "rain_per_step": 0.01, "steps": 3, "h0": 1.0, "expected_h": 1.03, "atol": 1e-12,
    },
    "horton_infiltration": {
        "source": "test_infiltration.jl:58-61,68-72", "f0": 1.977e-4, "fc": 3.272e-5, "k": 2.43e-3,
        "queries": [[0.0, 1.977e-4], [1.0e6, 3.272e-5]], "run": {"t0": 1e6, "T": 1000.0, "atol": 1e-10},
    },
    "swashes412": {
        "source": "test/swashes.jl:58-71,131-165 (dry-bed dam break: L=10, x0=5, hl=0.005, hr=0, g=9.81)",
        "t": 4.0,
        # x, analytic h, analytic u of the Ritter solution at t = 4 (evaluated by tests/cases.swashes412_analytic)
        "samples": [],
    },
}


def sha(arrs) -> str:
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_case(oracle, name, dtype):
    case, state = GOLDEN_CASES[name](dtype)
    o = oracle.OracleSim(case)
    o.set_state_interior(*state)
    mts = case.get("maximum_timestep", math.inf)
    dts = []
    for _ in range(NSTEPS):
        o.step(mts)
        dts.append(o.dt)
    full = o.get_state_full()
    inputs = list(state) + ([np.asarray(case["B"])] if isinstance(case.get("B"), np.ndarray) else [])
    return full[:, 2:-2, 2:-2].copy(), np.array(dts, dtype=np.float64), o.t, sha(inputs)


def export_inputs(dirname):
    """parameters + raw little-endian input arrays of every golden case, for tests/golden/dump_reference.jl"""
    d = Path(dirname)
    d.mkdir(parents=True, exist_ok=True)
    for name in GOLDEN_CASES:
        for dtype in ("float64", "float32"):
            case, state = GOLDEN_CASES[name](dtype)
            key = f"{name}.{dtype}"
            B = case.get("B")
            inf = case.get("infiltration")
            params = {"dtype": dtype, "nx": case["nx"], "ny": case["ny"], "gc": case.get("gc", 2), "bc": case.get("bc", "periodic"),
                      "extent": [float(x) for x in case.get("extent", [0.0, 1.0, 0.0, 1.0])],
                      "depth_cutoff": case.get("depth_cutoff", 1e-5), "kappa": case.get("kappa", 1e-5),
                      "theta": case.get("theta", 1.2), "cfl": case.get("cfl", 0.25), "stepper": case.get("stepper", "rk2"),
                      "recon": case.get("recon", "linear"), "friction": case.get("friction"),
                      "bottom_source": bool(case.get("bottom_source", False)), "rain": case.get("rain"),
                      "infiltration": None if inf is None else ("horton" if isinstance(inf, dict) else float(inf)),
                      "maximum_timestep": case.get("maximum_timestep"), "has_B": isinstance(B, np.ndarray), "nsteps": NSTEPS}
            (d / f"{key}.json").write_text(json.dumps(params, indent=1) + "\n")
            for nm, a in zip(("w", "hu", "hv"), state):
                np.ascontiguousarray(a, dtype=dtype).tofile(d / f"{key}.{nm}.bin")    # (ny, nx) C-order == Julia (nx, ny)
            if isinstance(B, np.ndarray):
                np.ascontiguousarray(B, dtype=dtype).tofile(d / f"{key}.B.bin")
    print("exported", len(list(d.glob("*.json"))), "cases to", d)


def import_reference(dirname):
    """packs the outputs of dump_reference.jl into tests/golden/reference_trajectories.npz"""
    d = Path(dirname)
    out = {}
    for pj in sorted(d.glob("*.json")):
        key = pj.name[:-5]
        p = json.loads(pj.read_text())
        st = np.fromfile(d / f"{key}.ref_state.bin", dtype=p["dtype"]).reshape(3, p["ny"], p["nx"])
        out[key + ".state"] = st
        out[key + ".dt"] = np.fromfile(d / f"{key}.ref_dt.bin", dtype=np.float64)
    np.savez_compressed(HERE / "reference_trajectories.npz", **out)
    print("wrote reference_trajectories.npz with", len(out) // 2, "cases")


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--export-inputs":
        return export_inputs(sys.argv[2])
    if len(sys.argv) == 3 and sys.argv[1] == "--import-reference":
        return import_reference(sys.argv[2])
    oracle = cases.load_oracle()
    out = {}
    meta = {"nsteps": NSTEPS, "cases": {}}
    for name in GOLDEN_CASES:
        for dtype in ("float64", "float32"):
            st, dts, t, h = run_case(oracle, name, dtype)
            key = f"{name}.{dtype}"
            out[key + ".state"] = st
            out[key + ".dt"] = dts
            meta["cases"][key] = {"t": t, "inputs_sha256": h, "shape": list(st.shape)}
    np.savez_compressed(HERE / "trajectories.npz", **out)
    (HERE / "trajectories.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")

    ka = json.loads(json.dumps(KNOWN_ANSWERS))
    xs = np.array([1.0, 3.0, 4.2, 4.8, 5.0, 5.5, 6.0, 6.5, 6.7, 9.0])
    h, u = cases.swashes412_analytic(xs, 4.0)
    ka["swashes412"]["samples"] = [[float(a), float(b), float(c)] for a, b, c in zip(xs, h, u)]
    (HERE / "reference_known_answers.json").write_text(json.dumps(ka, indent=1) + "\n")
    print("wrote", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
