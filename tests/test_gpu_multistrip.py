"""y-strip decomposition on the GPU: P strips (on P devices when the box has them, else all on
device 0 -- the peer pushes then go through the same code path with plain device pointers) must
reproduce the single-domain run BITWISE in the strict policies, and the oracle."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return pkg


def devices_for(nstrips):
    import torch
    n = torch.cuda.device_count()
    return list(range(nstrips)) if n >= nstrips else [0] * nstrips


CASES = {
    "lake_wall_rk2": lambda dt: cases.lake_case(n=96, ny=83, dtype=dt, level=0.0, wavelength=48.0),
    "dry_periodic_rk2": lambda dt: cases.dry_states_2d(dtype=dt, n=64),
    "terrain_wall_fe": lambda dt: cases.terrain_case(n=72, dtype=dt, level=0.0, wavelength=48.0, friction="bh2021", stepper="fe"),
    "urban_norecon": lambda dt: cases.urban_case(n=80, dtype=dt, level=0.5, wavelength=64.0),
    # periodic wrap-around WITH bathymetry: the first / last strips rebuild their ghost rows with the bathymetry of the far end
    "terrain_periodic_rk2": lambda dt: cases.terrain_case(n=72, dtype=dt, bc="periodic", level=0.0, wavelength=48.0, friction="bh2021"),
}


@pytest.mark.parametrize("nstrips", [2, 3])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_strips_match_single_domain_bitwise(P, oracle, name, dtype, nstrips):
    case, state = CASES[name](dtype)
    mts = case.get("maximum_timestep")
    N = 20
    single = P.simulator_from_case(case, strict=True)
    single.set_state_interior(*state)
    single.run_steps(N, mts)
    ref = single.get_state_interior()
    grp = P.multigpu.LocalStripGroup(case, state, devices_for(nstrips), strict=True)
    grp.run_steps(N, mts)
    got = grp.gather_interior()
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert grp.t == single.t and grp.dt == single.dt
    for s in grp.sims[1:]:
        assert s.t == single.t and s.dt == single.dt
    if case.get("friction") not in ("bsa2012", "fcg2016"):      # cbrt: last-ulp differences vs glibc
        o = oracle.OracleSim(case)
        o.set_state_interior(*state)
        for _ in range(N):
            o.step(mts if mts else float("inf"))
        for a, b in zip(got, o.get_state_interior()):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("name,nstrips", [("lake_wall_rk2", 4), ("dry_periodic_rk2", 3), ("urban_norecon", 2), ("terrain_wall_fe", 3),
                                          ("terrain_periodic_rk2", 3)])
def test_fast_strips_match_single_domain(P, name, nstrips, dtype):
    """fast arithmetic (the packed Float32 x2 kernels, the Float64 d1 kernels): a strip run executes exactly the same
    per-cell instruction sequence as the single-domain run, so the results are bitwise equal here too"""
    case, state = CASES[name](dtype)
    mts = case.get("maximum_timestep")
    single = P.simulator_from_case(case)
    single.set_state_interior(*state)
    single.run_steps(30, mts)
    grp = P.multigpu.LocalStripGroup(case, state, devices_for(nstrips))
    assert {s.kernel_family() for s in grp.sims} == {single.kernel_family()} <= {"fused-x2", "fused-d1"}
    grp.run_steps(30, mts)
    for a, b in zip(grp.gather_interior(), single.get_state_interior()):
        assert np.array_equal(a, b)
    assert grp.t == single.t and grp.dt == single.dt


def test_cost_balanced_strips_bitwise(P):
    """unequal strip heights (multigpu.partition_rows with a per-row cost) reproduce the single-domain run too"""
    case, state = CASES["lake_wall_rk2"]("float64")
    ny = case["ny"]
    cost = P.multigpu.row_cost_from_wet_fraction(np.linspace(0.0, 1.0, ny) ** 2)
    offs, sizes = P.multigpu.partition_rows(ny, 3, cost)
    assert sum(sizes) == ny and len(set(sizes)) > 1 and min(sizes) >= 4
    single = P.simulator_from_case(case, strict=True)
    single.set_state_interior(*state)
    single.run_steps(15)
    grp = P.multigpu.LocalStripGroup(case, state, devices_for(3), strict=True, row_cost=cost)
    assert [n for _, n in grp.rows] == sizes
    grp.run_steps(15)
    for a, b in zip(grp.gather_interior(), single.get_state_interior()):
        assert np.array_equal(a, b)
    assert grp.t == single.t


def test_partition_and_errors(P):
    off, sz = P.multigpu.partition_rows(83, 3)
    assert sz == [28, 28, 27] and off == [0, 28, 56]
    with pytest.raises(P.ArgumentError):
        P.multigpu.partition_rows(10, 4)
