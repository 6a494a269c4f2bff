import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import cases
P = cases.load_package()
case, state = cases.lake_case(n=1024, dtype="float32")
s = P.simulator_from_case(case); s.set_state_interior(*state)
for n in (0, 1, 2, 5):
    if n: s.run_steps(n)
    w, hu, hv = s.get_state_interior()
    Bc = cases.synthetic().cell_average(np.asarray(case["B"]))[2:-2, 2:-2]
    wet = (w - Bc) >= case["depth_cutoff"]
    nz = (hu != 0) | (hv != 0)
    print("after", n, "more steps: wet %.3f  nonzero-momentum %.3f  nonzero in dry cells %d  rows(64-col tiles) fully still: %.3f" % (
        wet.mean(), nz.mean(), int((nz & ~wet).sum()), 1 - nz.reshape(nz.shape[0], -1, 64).any(axis=2).mean()))
