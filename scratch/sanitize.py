"""small runs of every kernel family, to be executed under `compute-sanitizer --tool memcheck` (scratch/gpu_r02t.sh)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import cases
P = cases.load_package()
runs = []
for dtype in ("float32", "float64"):
    for bc in ("wall", "periodic", "neumann"):
        case, state = cases.terrain_case(n=70, dtype=dtype, bc=bc, level=0.0, wavelength=48.0, friction="bh2021")
        for strict in (False, True):
            s = P.simulator_from_case(case, strict=strict)
            s.set_state_interior(*state)
            s.run_steps(3, case.get("maximum_timestep"))
            s.sync()
            w = s.get_state_interior()[0]
            runs.append((dtype, bc, strict, s.kernel_family(), float(np.abs(w).max())))
    case, state = cases.urban_case(n=66, dtype=dtype, level=0.5, wavelength=48.0)
    s = P.simulator_from_case(case)
    s.set_state_interior(*state); s.run_steps(3, 1.0); s.sync(); runs.append((dtype, "urban", False, s.kernel_family(), 0.0))
    case, state = cases.terrain_case(n=64, dtype=dtype, bc="neumann", level=0.0, wavelength=48.0, friction="bh2021")
    grp = P.multigpu.LocalStripGroup(case, state, [0, 0, 0])
    grp.run_steps(3, 1.0); grp.sync(); runs.append((dtype, "neumann strips", False, grp.sims[0].kernel_family(), 0.0))
    case, state = cases.terrain_case(n=64, dtype=dtype, bc="periodic", level=0.0, wavelength=48.0, friction="bh2021")
    grp = P.multigpu.LocalStripGroup(case, state, [0, 0])
    grp.run_steps(3, 1.0); grp.sync(); runs.append((dtype, "periodic strips", False, grp.sims[0].kernel_family(), 0.0))
a = P.simulator_from_case(*[cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)[0]])
b = P.simulator_from_case(cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)[0], strict=True)
st = cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)[1]
a.set_state_interior(*st); b.set_state_interior(*st); a.run_steps(2); b.run_steps(2)
print(a.compare_states(b, tol=1e-5))
print(a.diagnostics())
for r in runs:
    print(r)
print("sanitize runs done:", len(runs))
