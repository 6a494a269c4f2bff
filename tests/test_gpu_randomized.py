"""Randomised configuration sweep on the GPU: seeded random combinations of grid shape (ragged, tiny, odd and even
widths), ghost cells, boundary condition, reconstruction, stepper, friction law, bathymetry, rain / infiltration
(incl. a Horton factor plane), thresholds and dtype, with partly dry initial states.

  * strict policy : the CUDA path (fused where the configuration allows it, per-operator kernels otherwise) must
                    reproduce the CPU oracle BIT FOR BIT after 8 steps (cbrt-based friction laws: 1e-13 / 1e-6);
  * fast policy   : on fully wet (well conditioned) states within the north_star tolerance -- this drives the packed
                    Float32 kernels through odd shapes, every BC / friction / source combination they support.
"""
import math

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


def random_case(seed: int, wet_only: bool = False, fused_only: bool = False):
    rng = np.random.default_rng(1000 + seed)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    dtype = pick(["float64", "float32"])
    gc = 2 if fused_only else pick([2, 2, 2, 1, 3])
    recon = "none" if gc == 1 else pick(["linear", "linear", "none"])
    bc = pick(["periodic", "wall"]) if fused_only else pick(["periodic", "wall", "wall", "neumann"])
    nx = int(rng.integers(4, 71))
    ny = int(rng.integers(4, 71))
    if seed % 7 == 0:
        nx, ny = pick([(4, 4), (5, 4), (4, 67), (130, 5), (61, 64), (124, 9)])
    dx, dy = float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.5, 2.0))
    cutoff = pick([1e-5, 1e-3, 8e-2])
    kappa = pick([cutoff, 1e-5, 1e-3])
    bkind = pick(["none", "const", "field", "field"])
    npd = np.dtype(dtype)
    Nx, Ny = nx + 2 * gc, ny + 2 * gc
    case = dict(dim=2, dtype=dtype, nx=nx, ny=ny, gc=gc, bc=bc, extent=[0.0, nx * dx, 0.0, ny * dy], recon=recon,
                theta=float(rng.uniform(1.0, 2.0)), stepper=pick(["rk2", "rk2", "fe"]), cfl=float(rng.uniform(0.08, 0.25)),
                depth_cutoff=cutoff, kappa=kappa, rho=pick([1.0, 1.0, 0.5]), g=9.81)
    if bkind == "const":
        case["B"] = float(rng.uniform(-1.0, 1.0))
        Bcell = np.full((ny, nx), case["B"])
    elif bkind == "field":
        x = np.arange(Nx + 1)[None, :] * dx
        y = np.arange(Ny + 1)[:, None] * dy
        B = (0.6 * np.sin(x / rng.uniform(3, 15) + rng.uniform(0, 6)) * np.cos(y / rng.uniform(3, 15) + rng.uniform(0, 6))
             + 0.02 * x + 0.15 * rng.standard_normal((Ny + 1, Nx + 1)))
        case["B"] = B.astype(npd)
        Bcell = cases.synthetic().cell_average(case["B"])[gc:-gc, gc:-gc]
    else:
        Bcell = np.zeros((ny, nx))
    if bkind != "none":
        case["bottom_source"] = True
    fr = pick([None, "bh2021", "bh2021", "bsa2012", "fcg2016"])
    if fr:
        case["friction"], case["Cz"] = fr, float(rng.uniform(1e-4, 5e-3))
    if rng.random() < 0.4:
        case["rain"] = pick([0.02, ([0.03, 0.001, 0.05], [0.0, 0.4, 1.1])])
    if rng.random() < 0.4:
        if rng.random() < 0.5:
            case["infiltration"] = float(rng.uniform(1e-6, 1e-5))
        else:
            case["infiltration"] = {"factor": rng.uniform(0.0, 2.0, size=(Ny, Nx))} if rng.random() < 0.6 else {}
        case["infil_first"] = bool(rng.random() < 0.5)
    # initial state: a smooth free surface cut by the terrain (dry where it is below the bathymetry)
    X = (np.arange(nx)[None, :] + 0.5) * dx
    Y = (np.arange(ny)[:, None] + 0.5) * dy
    level = float(np.quantile(Bcell, 0.55)) + 0.3 if bkind == "field" else float(np.max(Bcell)) + 0.5
    surf = level + 0.2 * np.sin(X / (0.31 * nx * dx) + 1.0) * np.cos(Y / (0.27 * ny * dy))
    if wet_only:
        surf = surf + (np.max(Bcell) - np.min(surf)) + 0.5
    w = np.maximum(surf, Bcell).astype(npd)
    h = w.astype(np.float64) - Bcell
    wet = h > 10 * cutoff
    hu = np.where(wet, 0.3 * h * np.sin(Y / (0.4 * ny * dy)), 0.0).astype(npd)
    hv = np.where(wet, -0.2 * h * np.cos(X / (0.5 * nx * dx)), 0.0).astype(npd)
    return case, (w, hu, hv)


def fused_possible(case):
    return case["gc"] == 2 and case["bc"] != "neumann"


@pytest.mark.parametrize("seed", range(48))
def test_random_config_strict_bit_exact(P, oracle, seed):
    case, state = random_case(seed)
    s = P.simulator_from_case(case, strict=True, fused=True)
    o = oracle.OracleSim(case)
    s.set_state_interior(*state)
    o.set_state_interior(*state)
    gc = case["gc"]
    cbrt = case.get("friction") in ("bsa2012", "fcg2016")
    for n in range(8):
        s.step(0.05)
        o.step(0.05)
        if cbrt:
            assert s.dt == pytest.approx(o.dt, rel=1e-12)
        else:
            assert s.dt == o.dt and s.t == o.t, f"seed {seed} step {n}: dt {s.dt} vs {o.dt}"
    a = s.get_state_full()[:, gc:-gc, gc:-gc]
    b = o.get_state_full()[:, gc:-gc, gc:-gc]
    assert np.all(np.isfinite(b))
    if cbrt:
        assert max(cases.field_errors(a, b)) <= (1e-12 if case["dtype"] == "float64" else 1e-5)
    else:
        assert np.array_equal(a, b), f"seed {seed}: {case['nx']}x{case['ny']} gc={gc} {case['bc']} {case['recon']} max diff {np.max(np.abs(a - b))}"


@pytest.mark.parametrize("seed", range(32))
def test_random_config_fast_within_tolerance(P, oracle, seed):
    case, state = random_case(seed, wet_only=True, fused_only=True)
    s = P.simulator_from_case(case, strict=False, fused=True)
    o = oracle.OracleSim(case)
    s.set_state_interior(*state)
    o.set_state_interior(*state)
    for _ in range(8):
        s.step(0.05)
        o.step(0.05)
    f64 = case["dtype"] == "float64"
    assert s.t == pytest.approx(o.t, rel=1e-12 if f64 else 1e-5)
    a = s.get_state_full()[:, 2:-2, 2:-2]
    b = o.get_state_full()[:, 2:-2, 2:-2]
    err = max(cases.field_errors(a, b))
    assert err <= (1e-12 if f64 else 1e-5), f"seed {seed}: {case['nx']}x{case['ny']} {case['bc']} {case['recon']} {case.get('friction')} err {err:.2e}"
