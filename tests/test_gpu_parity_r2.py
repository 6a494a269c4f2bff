"""Round-2 GPU parity tests: the performance path pinned per cell at FULL size against the strict (bit-faithful)
path, partially-dry random configurations with the oracle's own sensitivity as the yardstick, the SWASHES cases
the reference's tests run besides 4.1.2 (4.1.1: test_swashes.jl:356; 4.2.2.b: test/swashes.jl:106-119,260-289), the
kernel-family report, the device-side state comparison and the error returns of simulate_to_time.

Measured tables of these comparisons: profiles/r02_parity.md (scratch/parity_report.py)."""
import math
import os

import numpy as np
import pytest

import cases
from test_gpu_randomized import random_case

pytestmark = pytest.mark.gpu

F32_TOL, F64_TOL = 1e-5, 1e-12


@pytest.fixture(scope="module")
def P(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return pkg


# ------------------------------------------------------------------------------------------------------------------
def test_kernel_family_report(P):
    """swe_b200_kernel_family: which kernels serve a configuration is visible, not silent"""
    lake32, _ = cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)
    lake64, _ = cases.lake_case(n=64, ny=48, dtype="float64", wavelength=32.0)
    fam = lambda case, **kw: P.simulator_from_case(case, **kw).kernel_family()
    assert fam(lake32) == "fused-x2"                                # the bench line's kernels
    assert fam(lake32, strict=True) == "fused-generic"
    assert fam(lake64) in ("fused-d1", "fused-generic")
    assert fam(lake64, strict=True) == "fused-generic"
    assert fam(lake32, fused=False) == "per-operator"
    assert fam(dict(lake32, nx=63, B=np.asarray(lake32["B"])[:, :-1].copy(), extent=[0.0, 63.0, 0.0, 48.0])) == "fused-generic"   # odd pitch
    pure, _ = cases.gaussian_pure_2d()
    assert fam(pure) == "per-operator"
    urban, _ = cases.urban_case(n=64, dtype="float32", level=1.0, wavelength=32.0)
    assert fam(urban) == "fused-x2"
    assert fam(dict(urban, bc="neumann")) == "fused-x2"             # NeumannBC (examples/urban.jl:125) on the packed kernels
    assert fam(dict(lake64, bc="neumann")) in ("fused-d1", "fused-generic")
    assert fam(dict(urban, bc="neumann"), fused=False) == "per-operator"


def test_compare_states_matches_numpy(P):
    case, state = cases.lake_case(n=100, ny=77, dtype="float32", wavelength=48.0)
    a = P.simulator_from_case(case, strict=False)
    b = P.simulator_from_case(case, strict=True)
    for s in (a, b):
        s.set_state_interior(*state)
        s.run_steps(30)
    c = a.compare_states(b, tol=1e-6)
    x, y = a.get_state_full()[:, 2:-2, 2:-2], b.get_state_full()[:, 2:-2, 2:-2]
    want = cases.field_errors(x, y)
    assert c["rel_linf"] == pytest.approx(want, rel=1e-12)
    Bc = cases.synthetic().cell_average(np.asarray(case["B"]))[2:-2, 2:-2].astype(np.float32)
    wet = lambda w: (w - Bc).astype(np.float64) >= case["depth_cutoff"]
    assert c["mask_mismatch"] == int(np.sum(wet(x[0]) != wet(y[0])))
    assert c["wet"] == int(np.sum(wet(x[0]))) and c["wet_other"] == int(np.sum(wet(y[0]))) and c["nan_cells"] == 0
    scale = [np.max(np.abs(y[0])), max(np.max(np.abs(y[1])), np.max(np.abs(y[2])))]
    over = [int(np.sum(np.abs(x[k].astype(np.float64) - y[k]) > 1e-6 * scale[min(k, 1)])) for k in range(3)]
    assert c["cells_over_tol"] == over
    other, _ = cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)
    with pytest.raises(P.ArgumentError):
        a.compare_states(P.simulator_from_case(other))


# ------------------------------------------------------------------------------------------------------------------
# FULL SIZE: BASELINE configs[4] (the bench line) per cell, fast vs strict, with the strict path's own +-1 ulp
# sensitivity as the yardstick (the strict path is bit-exact against the oracle wherever the oracle can run).
# ------------------------------------------------------------------------------------------------------------------
def _full_size(P, build, n, dtype, checkpoints, tol):
    import torch
    need = 3 * 3 * 3 * (n + 4) ** 2 * np.dtype(dtype).itemsize * 1.3
    if torch.cuda.mem_get_info()[0] < need:
        pytest.skip("not enough free device memory for three full-size handles")
    case, state = build(n)
    mts = case.get("maximum_timestep")
    fast = P.simulator_from_case(case, strict=False)
    strict = P.simulator_from_case(case, strict=True)
    pert = P.simulator_from_case(case, strict=True)
    fast.set_state_interior(*state)
    strict.set_state_interior(*state)
    pert.set_state_interior(*cases.perturb_ulp(state, dtype, seed=1))
    del state
    done, rows = 0, []
    for n_steps in checkpoints:
        for s in (fast, strict, pert):
            s.run_steps(n_steps - done, mts)
        done = n_steps
        c, cp = fast.compare_states(strict, tol=tol), pert.compare_states(strict, tol=tol)
        rows.append((n_steps, c, cp))
        assert c["nan_cells"] == 0
        # within the tolerance, or within 4x what a +-1 ulp input perturbation does to the reference arithmetic itself
        # (the two momenta share one scale and one yardstick: the sensitivity of a single seed is heavy-tailed per field)
        assert c["rel_linf"][0] <= max(tol, 4 * cp["rel_linf"][0]), (n_steps, c["rel_linf"], cp["rel_linf"])
        sens_m = max(cp["rel_linf"][1], cp["rel_linf"][2])
        for k in (1, 2):
            assert c["rel_linf"][k] <= max(tol, 4 * sens_m), (n_steps, k, c["rel_linf"], cp["rel_linf"])
        # identical wet/dry masks up to what the +-1 ulp perturbation flips
        assert c["mask_mismatch"] <= 2 * cp["mask_mismatch"] + 4, (c["mask_mismatch"], cp["mask_mismatch"])
        assert abs(fast.dt - strict.dt) <= 1e-6 * strict.dt
    return fast.kernel_family(), rows


def test_full_size_config5_fast_vs_strict_per_cell(P):
    n = int(os.environ.get("SWE_B200_FULL_N", 16384))
    fam, rows = _full_size(P, lambda m: cases.lake_case(n=m, dtype="float32", level=0.0), n, "float32", (10, 20), F32_TOL)
    assert fam == "fused-x2"
    for n_steps, c, cp in rows:
        # the free surface of the bench workload is well conditioned: the north_star tolerance holds for EVERY cell
        assert c["rel_linf"][0] <= F32_TOL and c["cells_over_tol"][0] == 0 and c["mask_mismatch"] <= 1
        print(f"\nconfig 5 {n}^2 N={n_steps}: fast vs strict {['%.1e' % e for e in c['rel_linf']]} (cells > 1e-5: {c['cells_over_tol']}), "
              f"strict 1-ulp sens {['%.1e' % e for e in cp['rel_linf']]} (cells > 1e-5: {cp['cells_over_tol']}), "
              f"mask mismatches {c['mask_mismatch']} / {cp['mask_mismatch']} of {c['wet_other']} wet")


def test_full_size_lake_at_rest_is_preserved(P):
    """size-independent property at BASELINE's full size (test_lake_at_rest_2d.jl:97-98): the bench workload's terrain fully
    submerged under a flat free surface (w = 12) is a steady state of the well-balanced scheme.  On the packed Float32
    kernels at 16384^2, 20 steps: the free surface does not move by a single ulp, the momenta stay at the Float32 round-off
    of the hydrostatic balance (the CPU oracle in Float32 on a 512^2 instance: max|hu| = 2.1e-5), the volume is conserved."""
    import torch
    n = int(os.environ.get("SWE_B200_FULL_N", 16384))
    if torch.cuda.mem_get_info()[0] < 20e9 * (n / 16384) ** 2:
        pytest.skip("not enough free device memory")
    case, _ = P.workloads.lake_case(n=n, dtype="float32", level=12.0)
    s = P.simulator_from_case(case)
    assert s.kernel_family() == "fused-x2"
    w = np.full((n, n), 12.0, dtype=np.float32)
    z = np.zeros_like(w)
    s.set_state_interior(w, z, z)
    del w, z
    d0 = s.diagnostics()
    s.run_steps(20)
    d1 = s.diagnostics()
    assert d1["wet_cells"] == d0["wet_cells"] == n * n
    assert d0["sum_w"] == 12.0 * n * n
    assert abs(d1["sum_w"] - d0["sum_w"]) <= 1e-8 * d0["sum_w"]       # (the oracle keeps w bitwise; allow single-ulp moves)
    assert abs(d1["max_h"] - d0["max_h"]) <= 2e-6 and abs(d1["min_h"] - d0["min_h"]) <= 2e-6
    assert d1["max_momentum"] <= 1e-4, d1                             # 4-5 x the oracle's Float32 round-off level


@pytest.mark.parametrize("cfg,n", [(3, 2048), (4, 4096)])
def test_large_configs_fast_vs_strict_per_cell(P, cfg, n):
    build = {3: lambda m: cases.terrain_case(n=m, dtype="float32", level=1.0),
             4: lambda m: cases.urban_case(n=m, dtype="float32", level=1.0)}[cfg]
    fam, rows = _full_size(P, build, n, "float32", (10, 20), F32_TOL)
    assert fam == "fused-x2"


# ------------------------------------------------------------------------------------------------------------------
# partially dry random configurations on the fast path (the fully wet ones are in test_gpu_randomized.py)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(100, 124))
def test_random_partially_dry_fast_within_sensitivity(P, oracle, seed):
    case, state = random_case(seed, wet_only=False, fused_only=True)
    dtype = case["dtype"]
    tol = F64_TOL if dtype == "float64" else F32_TOL

    def run_oracle(st):
        o = oracle.OracleSim(case)
        o.set_state_interior(*st)
        for _ in range(8):
            o.step(0.05)
        return o.get_state_full()[:, 2:-2, 2:-2], o.t

    b, t_ref = run_oracle(state)
    perts = [run_oracle(cases.perturb_ulp(state, dtype, seed=sd))[0] for sd in (1, 2, 3)]
    sens = np.max([cases.field_errors(pp, b) for pp in perts], axis=0)
    s = P.simulator_from_case(case, strict=False, fused=True)
    s.set_state_interior(*state)
    for _ in range(8):
        s.step(0.05)
    a = s.get_state_full()[:, 2:-2, 2:-2]
    err = cases.field_errors(a, b)
    for e, sn, nm in zip(err, sens, ("w", "hu", "hv")):
        assert e <= max(tol, 4 * sn), f"seed {seed} {nm}: {e:.2e} (1-ulp sensitivity {sn:.2e}) {case['nx']}x{case['ny']} {case['bc']} {case['recon']} {case.get('friction')}"
    B = case.get("B")
    if isinstance(B, np.ndarray):
        Bc = cases.synthetic().cell_average(B)[2:-2, 2:-2]
    else:
        Bc = float(B or 0.0)
    wet = lambda w: (w.astype(np.float64) - Bc) >= case["depth_cutoff"]
    mism = int(np.sum(wet(a[0]) != wet(b[0])))
    mism_ref = max(int(np.sum(wet(pp[0]) != wet(b[0]))) for pp in perts)
    assert mism <= 2 * mism_ref + 2, f"seed {seed}: wet/dry masks differ in {mism} cells (perturbed oracle: {mism_ref})"
    assert s.t == pytest.approx(t_ref, rel=max(10 * tol, 1e-6))


# ------------------------------------------------------------------------------------------------------------------
# SWASHES (test/swashes.jl, test/test_swashes.jl)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("strict", [False, True])
def test_swashes_411_wet_dam_break(P, oracle, strict):
    """test_swashes.jl:356 compare_swashes_in2d(swashes411, nx, 6.0): x-run, y-run and 1D run agree (bitwise in the
    strict policy) and the result is Stoker's solution (swashes.jl:131-156) to discretisation accuracy."""
    nx, T = 256, 6.0
    sims = {}
    for layout in ("x", "y"):
        case, st = cases.swashes41x("411", nx, layout)
        s = P.simulator_from_case(case, strict=strict)
        s.set_state_interior(*st)
        P.simulate_to_time(s, T)
        sims[layout] = s
    wx, hux, hvx = sims["x"].get_state_interior()
    wy, huy, hvy = sims["y"].get_state_interior()
    assert np.all(hvx == 0) and np.all(huy == 0)
    assert np.array_equal(wy[:, 1], wx[1, :]) and np.array_equal(hvy[:, 1], hux[1, :])
    assert sims["x"].dt == sims["y"].dt and sims["x"].t == T
    case1, st1 = cases.swashes41x("411", nx, "1d")
    o = oracle.OracleSim(case1)
    o.set_state_interior(*st1)
    o.simulate_to_time(T)
    h1, hu1 = o.get_state_interior()
    if strict:
        assert np.array_equal(h1, wx[1, :]) and np.array_equal(hu1, hux[1, :]) and o.dt == sims["x"].dt
    else:
        assert cases.rel_linf(wx[1, :], h1) <= 1e-9
    xc = cases.synthetic().cell_centers(nx, 2, 0.0, 10.0)
    h_ref, u_ref = cases.swashes411_analytic(xc, T)
    # the oracle's own error against the analytic solution at this resolution: mean 7.7e-6, max 5.6e-4 (at the shock)
    assert np.mean(np.abs(wx[1, :] - h_ref)) < 1.2e-5
    away = np.abs(xc - (5.0 + T * (2 * 0.1578324867 ** 2 * (np.sqrt(9.81 * 0.005) - 0.1578324867)) / (0.1578324867 ** 2 - 9.81 * 0.001))) > 0.3
    assert np.max(np.abs(wx[1, :] - h_ref)[away]) < 1.5e-4


def test_swashes_411_2d_config2_wet_variant(P):
    """SURVEY config 2, wet variant: the 2D RK2 run on an n x n grid is x-invariant in y and matches Stoker"""
    n, T = 256, 6.0
    case, st = cases.swashes_dambreak_2d(n=n, dtype="float64", kind="411")
    s = P.simulator_from_case(case)
    s.set_state_interior(*st)
    P.simulate_to_time(s, T)
    w, hu, hv = s.get_state_interior()
    assert s.t == T and np.all(hv == 0) and np.array_equal(w[0], w[n // 2]) and np.array_equal(hu[3], hu[-1])
    xc = cases.synthetic().cell_centers(n, 2, 0.0, 10.0)
    h_ref, _ = cases.swashes411_analytic(xc, T)
    assert np.mean(np.abs(w[n // 2] - h_ref)) < 1.2e-5


@pytest.mark.parametrize("strict", [False, True])
def test_swashes_421_planar_surface_in_parabola_1d(P, oracle, strict):
    """test/swashes.jl:75-86, 183-229: 1D, BottomTopography1D, moving shorelines, through the 1D embedding of the 2D kernels;
    a tenth of the period (see cases.swashes421): strict == the oracle's genuine 1D path to 1e-13, analytic to 1.5x the
    oracle's own error (4.8e-5)"""
    case, st, period = cases.swashes421(256)
    s = P.simulator_from_case(case, strict=strict)
    o = oracle.OracleSim(case)
    for x in (s, o):
        x.set_state_interior(*st)
    T = 0.1 * period
    ns = P.simulate_to_time(s, T)
    no = o.simulate_to_time(T)
    w, hu = s.get_state_interior()
    wo, huo = o.get_state_interior()
    assert s.t == T and (ns == no or not strict)
    assert cases.rel_linf(w, wo) <= (1e-13 if strict else 1e-9)
    _, ref, _ = cases.swashes421(256, t=T)
    assert np.mean(np.abs(w - ref[0])) < 1.5 * 4.8e-5


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_swashes_422b_planar_surface_in_paraboloid(P, dtype):
    """test/swashes.jl:106-119, 260-289: a planar free surface rotating in a paraboloid, moving shoreline, one rotation.
    The reference never asserts against the analytic solution; the bound is 1.5x the CPU oracle's own error at this
    resolution (Float64, 100^2, one rotation: mean|w - w_ref| = 8.8e-4, mean|hv - hv_ref| = 9.0e-4)."""
    case, st, period = cases.swashes422b(100, dtype=dtype)
    s = P.simulator_from_case(case)
    s.set_state_interior(*st)
    v0 = s.diagnostics()["sum_h"]
    nsteps = P.simulate_to_time(s, period)
    w, hu, hv = s.get_state_interior()
    assert s.t == pytest.approx(period, rel=1e-6) and nsteps > 1000 and np.all(np.isfinite(w))
    _, ref, _ = cases.swashes422b(100, t=period)
    assert np.mean(np.abs(w - ref[0])) < 1.5 * 8.8e-4
    assert np.mean(np.abs(hv - ref[2])) < 1.5 * 9.0e-4 and np.mean(np.abs(hu - ref[1])) < 1.5 * 3.3e-4
    # wall BC: mass is conserved up to round-off and the dry clean-up's w = max(w, B_cell) (timestepper.jl:41-51), which
    # creates water in cells that dip below the bed: 3e-4 relative in Float32 over the ~64 000 steps of one rotation
    assert s.diagnostics()["sum_h"] == pytest.approx(v0, rel=1e-10 if dtype == "float64" else 1e-3)


# ------------------------------------------------------------------------------------------------------------------
# NeumannBC on the fused path (bc.jl:49-60; the BC of examples/terrain.jl:85 and examples/urban.jl:125)
# ------------------------------------------------------------------------------------------------------------------
NEUMANN_CASES = {
    "terrain_rk2": lambda dt: cases.terrain_case(n=96, dtype=dt, bc="neumann", level=0.0, wavelength=64.0, friction="bh2021"),
    "terrain_fe_dry": lambda dt: cases.terrain_case(n=72, dtype=dt, bc="neumann", wavelength=48.0, stepper="fe", film=0.12),
    "urban_norecon": lambda dt: cases.urban_case(n=80, dtype=dt, bc="neumann", level=0.5, wavelength=64.0),
    "lake_bh2021": lambda dt: dict(cases.lake_case(n=100, ny=77, dtype=dt, wavelength=48.0)[0], bc="neumann"),
}


def _neumann(name, dtype):
    got = NEUMANN_CASES[name](dtype)
    if isinstance(got, dict):
        return got, cases.lake_case(n=100, ny=77, dtype=dtype, wavelength=48.0)[1]
    return got


def _no_corners(a, gc=2):
    m = np.ones(a.shape[1:], dtype=bool)
    m[:gc, :gc] = m[:gc, -gc:] = m[-gc:, :gc] = m[-gc:, -gc:] = False
    return a[:, m]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(NEUMANN_CASES))
def test_neumann_fused_strict_bit_exact(P, oracle, name, dtype):
    """interior AND ghost cells (the self-evolving high-side ghosts included), every step's dt, fused == per-operator"""
    case, state = _neumann(name, dtype)
    mts = case.get("maximum_timestep", math.inf)
    s = P.simulator_from_case(case, strict=True, fused=True)
    u = P.simulator_from_case(case, strict=True, fused=False)
    o = oracle.OracleSim(case)
    assert s.kernel_family() == "fused-generic" and u.kernel_family() == "per-operator"
    for x in (s, u, o):
        x.set_state_interior(*state)
    cbrt = case.get("friction") in ("bsa2012", "fcg2016")
    for n in range(25):
        for x in (s, u, o):
            x.step(mts)
        if cbrt:
            assert s.dt == pytest.approx(o.dt, rel=1e-12)
        else:
            assert s.dt == o.dt == u.dt, f"step {n}: dt {s.dt} vs {o.dt}"
    a, b, c = (_no_corners(x.get_state_full()) for x in (s, o, u))
    if cbrt:
        assert max(cases.field_errors(a, b)) <= (1e-12 if dtype == "float64" else 1e-5)
    else:
        assert np.array_equal(a, b), f"max diff {np.max(np.abs(a - b))}"
        assert np.array_equal(a, c)
    assert np.any(s.get_state_full()[:, 2:-2, -1] != 0)            # the outermost high ghost column holds evolved data


@pytest.mark.parametrize("dtype,tol", [("float64", F64_TOL), ("float32", F32_TOL)])
@pytest.mark.parametrize("name", ["terrain_rk2", "lake_bh2021"])
def test_neumann_fused_fast_within_tolerance(P, oracle, name, dtype, tol):
    """fast policies under NeumannBC (x2 / d1 kernels).  The image of the top rows in the low-y ghosts and the
    self-evolving high-side ghosts put discontinuities on the boundary, so the lake case is ill-conditioned from the
    first steps (the oracle's own +-1 ulp sensitivity after 20 steps is 9e-5 in Float32; 3e-7 under WallBC): the bound is
    max(tolerance, 4 x that sensitivity) like in test_gpu_parity.py."""
    case, state = _neumann(name, dtype)
    mts = case.get("maximum_timestep", math.inf)

    def run_oracle(st):
        o = oracle.OracleSim(case)
        o.set_state_interior(*st)
        for _ in range(20):
            o.step(mts)
        return o.get_state_full()[:, 2:-2, 2:-2], o.t

    b, t_ref = run_oracle(state)
    # four seeds: the sensitivity is heavy-tailed (one dry-state switch flips or it does not: seeds 1-3 give 1e-5, seed 4 9e-5)
    sens = np.max([cases.field_errors(run_oracle(cases.perturb_ulp(state, dtype, seed=sd))[0], b) for sd in (1, 2, 3, 4)], axis=0)
    s = P.simulator_from_case(case)
    s.set_state_interior(*state)
    for _ in range(20):
        s.step(mts)
    assert s.kernel_family() == ("fused-d1" if dtype == "float64" else "fused-x2")
    err = cases.field_errors(s.get_state_full()[:, 2:-2, 2:-2], b)
    for e, sn in zip(err, sens):
        assert e <= max(tol, 4 * sn), (err, list(sens))
    assert s.t == pytest.approx(t_ref, rel=max(10 * tol, 1e-6, 4 * float(np.max(sens))))


def test_close_peers_then_destroy(P):
    """collective tear-down order of a strip group: unmap the peers first (swe_b200_close_peers), then destroy"""
    case, state = cases.lake_case(n=64, ny=48, dtype="float32", wavelength=32.0)
    grp = P.multigpu.LocalStripGroup(case, state, [0, 0])
    grp.run_steps(5)
    ref = grp.gather_interior()
    L = P.api._lib.lib()
    for s in grp.sims:
        s._check(L.swe_b200_close_peers(s._h))
        s._check(L.swe_b200_close_peers(s._h))          # idempotent
    parts = [s.get_state_interior() for s in grp.sims]  # the strips' own data stay readable
    got = [np.concatenate([p[k] for p in parts], axis=0) for k in range(3)]
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    for s in grp.sims:
        s.close()


@pytest.mark.parametrize("strict,dtype", [(True, "float64"), (True, "float32"), (False, "float32"), (False, "float64")])
@pytest.mark.parametrize("name,nstrips", [("terrain_rk2", 3), ("terrain_fe_dry", 2), ("urban_norecon", 4), ("lake_bh2021", 2)])
def test_neumann_strips_match_single_domain_bitwise(P, name, nstrips, strict, dtype):
    """NeumannBC on y-strips: the low-y ghosts of the first strip are the image of the LAST strip's top rows (a
    one-directional wrap-around push), the high-y ghosts evolve on the last strip, the high-x ghost column on every strip"""
    import torch
    case, state = _neumann(name, dtype)
    mts = case.get("maximum_timestep")
    single = P.simulator_from_case(case, strict=strict)
    single.set_state_interior(*state)
    single.run_steps(20, mts)
    ref = single.get_state_interior()
    nd = torch.cuda.device_count()
    devs = list(range(nstrips)) if nd >= nstrips else [0] * nstrips
    grp = P.multigpu.LocalStripGroup(case, state, devs, strict=strict)
    grp.run_steps(20, mts)
    for a, b in zip(grp.gather_interior(), ref):
        assert np.array_equal(a, b), float(np.max(np.abs(a - b)))
    assert grp.t == single.t and grp.dt == single.dt


# ------------------------------------------------------------------------------------------------------------------
def test_simulate_to_time_reports_stalled_time(P):
    """dt == 0 forever (here: a wave speed so large that t + dt == t) must come back as an error code, not hang the
    caller: the reference's loop (simulator.jl:199-210) would never end"""
    case, st = cases.dry_states_2d(n=32)
    case["t0"] = 1.0
    s = P.simulator_from_case(case)
    w = st[0].copy()
    w[5, 5] = 1e30
    s.set_state_interior(w, st[1], st[2])
    with pytest.raises(P.DomainError, match="does not advance"):
        P.simulate_to_time(s, 1.05)


def test_enumerate_grid_without_simulator(P):
    """the kernel surface binds (backend, grid), not a simulator: swe_b200_enumerate_grid"""
    be = P.make_b200_backend(np.float64)
    grid = P.CartesianGrid(5, 4, gc=2, boundary=P.WallBC())
    seen, extra = [], []
    P.for_each_cell(lambda I, y: (seen.append(I), extra.append(y)), be, grid, "y1")
    assert len(seen) == 9 * 8 and seen[0] == (1, 1) and seen[-1] == (9, 8) and set(extra) == {"y1"}
    tri = []
    P.for_each_inner_cell(lambda l, m, r, a, b: tri.append((l, m, r, a + b)), be, grid, P.XDIR, 1, 2, ghostcells=1)
    assert len(tri) == (9 - 2) * 4 and tri[0] == ((1, 3), (2, 3), (3, 3), 3)
    g = []
    P.for_each_ghost_cell(lambda I: g.append(I), be, grid, P.YDIR)
    assert len(g) == 5 * 2 and g[0] == (3, 1)
