"""Shared test cases: parameter sets of the reference's own tests / examples and of
BASELINE.json's configs, as plain dicts understood by both oracle.OracleSim and the
product's Simulator (volumefluxes.jl_b200.api.simulator_from_case).

Arrays are numpy C-order (Ny, Nx) == the reference's column-major (Nx, Ny).
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def load_package():
    """Import the product package (its directory name contains a dot, so use importlib)."""
    name = "volumefluxes_jl_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "volumefluxes.jl_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_oracle():
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    from oracle import oracle  # noqa
    return oracle


def synthetic():
    return load_package().synthetic


# ------------------------------------------------------------------------------------------
# SWASHES 4.1.1 / 4.1.2 dam breaks (test/swashes.jl:44-71,153-181; test/test_swashes.jl:218-258)
# ------------------------------------------------------------------------------------------
def swashes41x(kind="412", nx=512, layout="x", dtype="float64", stepper="fe", ny=3):
    L, x0, hl = 10.0, 5.0, 0.005
    hr = 0.0 if kind == "412" else 0.001
    syn = synthetic()
    xc = syn.cell_centers(nx, 2, 0.0, L)
    h = np.where(xc < x0, hl, hr)
    base = dict(dtype=dtype, gc=2, bc="wall", depth_cutoff=1e-7, kappa=1e-7, theta=1.2, cfl=0.1,
                stepper=stepper, recon="linear")
    if layout == "1d":
        case = dict(base, dim=1, nx=nx, extent=[0.0, L])
        state = (h.copy(), np.zeros(nx))
    elif layout == "x":
        case = dict(base, dim=2, nx=nx, ny=ny, extent=[0.0, L, 0.0, 5.0])
        w = np.tile(h[None, :], (ny, 1))
        state = (w, np.zeros_like(w), np.zeros_like(w))
    else:
        case = dict(base, dim=2, nx=ny, ny=nx, extent=[0.0, 5.0, 0.0, L])
        w = np.tile(h[:, None], (1, ny))
        state = (w, np.zeros_like(w), np.zeros_like(w))
    return case, state


def swashes_dambreak_2d(n=1024, dtype="float64", kind="412"):
    """BASELINE config 2: 2D SWASHES dam break on an n x n grid, RK2."""
    case, state = swashes41x(kind=kind, nx=n, layout="x", dtype=dtype, stepper="rk2", ny=n)
    case["extent"] = [0.0, 10.0, 0.0, 10.0]
    return case, state


def swashes412_analytic(x, t, g=9.81, hl=0.005, x0=5.0):
    """test/swashes.jl:131-165 (dry-bed dam break, Ritter solution)."""
    xA = x0 - t * np.sqrt(g * hl)
    xB = x0 + 2 * t * np.sqrt(g * hl)
    h = np.where(x < xA, hl, np.where(x < xB, (4.0 / (9.0 * g)) * (np.sqrt(g * hl) - (x - x0) / (2 * t)) ** 2, 0.0))
    u = np.where(x < xA, 0.0, np.where(x < xB, (2.0 / 3.0) * ((x - x0) / t + np.sqrt(g * hl)), 0.0))
    return h, u


def swashes411_analytic(x, t, g=9.81, hl=0.005, hr=0.001, cm=0.1578324867, x0=5.0):
    """test/swashes.jl:131-156 (dam break on a wet domain, Stoker's solution; cm from swashes.jl:54)."""
    c = np.sqrt(g * hl)
    xA = x0 - t * c
    xB = x0 + t * (2 * c - 3 * cm)
    xC = x0 + t * (2 * cm ** 2 * (c - cm)) / (cm ** 2 - g * hr)
    rar_h = (4.0 / (9.0 * g)) * (c - (x - x0) / (2 * t)) ** 2
    rar_u = (2.0 / 3.0) * ((x - x0) / t + c)
    h = np.where(x < xA, hl, np.where(x < xB, rar_h, np.where(x < xC, cm ** 2 / g, hr)))
    u = np.where(x < xA, 0.0, np.where(x < xB, rar_u, np.where(x < xC, 2 * (c - cm), 0.0)))
    return h, u


def swashes422b(n=100, dtype="float64", t=0.0, a=1.0, L=4.0, g=9.81, h0=0.1, eta=0.5, stepper="rk2"):
    """SWASHES 4.2.2.b, 2D planar surface in a paraboloid (test/swashes.jl:106-119, 260-289, 303-308): bathymetry at
    the corners, the analytic state at time t on the interior (w = max(plane, B_cell), hu = u h, hv = v h).
    Returns (case, state, rotation period 2 pi / omega)."""
    syn = synthetic()
    omega = np.sqrt(2 * g * h0) / a
    xf = syn.cell_faces(n, 2, 0.0, L, interior=False)
    X, Y = np.meshgrid(xf, xf)
    B = -h0 * (1 - ((X - L / 2) ** 2 + (Y - L / 2) ** 2) / a ** 2)
    xc = syn.cell_centers(n, 2, 0.0, L)
    XC, YC = np.meshgrid(xc, xc)
    Bc = syn.cell_average(B)[2:-2, 2:-2]
    cs, sn = np.cos(omega * t), np.sin(omega * t)
    w = np.maximum((eta * h0 / a ** 2) * (2 * (XC - L / 2) * cs + 2 * (YC - L / 2) * sn - eta), Bc)
    h = w - Bc
    state = (w, -eta * omega * sn * h, eta * omega * cs * h)
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc="wall", extent=[0.0, L, 0.0, L], depth_cutoff=1e-5,
                kappa=1e-5, theta=1.2, cfl=0.25, stepper=stepper, recon="linear", B=B.astype(dtype), bottom_source=True, g=g)
    return case, tuple(np.ascontiguousarray(s, dtype=dtype) for s in state), 2 * np.pi / omega


# ------------------------------------------------------------------------------------------
# lake at rest (test/test_lake_at_rest_2d.jl:28-119)
# ------------------------------------------------------------------------------------------
def lake_at_rest_2d(which="step", dtype="float64", stepper="fe"):
    syn = synthetic()
    if which == "step":
        n, ext = 64, [0.0, 10.0, 0.0, 10.0]
        xf = syn.cell_faces(n, 2, ext[0], ext[1], interior=False)
        yf = syn.cell_faces(n, 2, ext[2], ext[3], interior=False)
        B = np.where((xf[None, :] < 5.0) & (yf[:, None] < 5.0), 0.45, 0.55)
    else:
        n, ext = 124, [-2 * np.pi, 2 * np.pi, -2 * np.pi, 2 * np.pi]
        xf = syn.cell_faces(n, 2, ext[0], ext[1], interior=False)
        yf = syn.cell_faces(n, 2, ext[2], ext[3], interior=False)
        X, Y = xf[None, :], yf[:, None]
        B = np.cos(X + Y) - 0.5 - 1.5 * (np.sqrt(X ** 2 + Y ** 2) < 1.0)
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc="wall", extent=ext, theta=2.0, cfl=0.2,
                stepper=stepper, recon="linear", B=B, bottom_source=True,
                depth_cutoff=1e-5, kappa=1e-5)
    w = np.full((n, n), 0.7)
    return case, (w, np.zeros_like(w), np.zeros_like(w))


# ------------------------------------------------------------------------------------------
# test_shallow_water_2d.jl (Pure, Gaussian) and test_shallow_water_dry_states_2d.jl
# ------------------------------------------------------------------------------------------
def gaussian_pure_2d(dtype="float64", nx=256, ny=32):
    syn = synthetic()
    x = syn.cell_centers(nx, 2, 0.0, 1.0)
    y = syn.cell_centers(ny, 2, 0.0, 1.0)
    r2 = (x[None, :] - 0.5) ** 2 + (y[:, None] - 0.5) ** 2
    w = np.exp(-(r2 / 0.01)) + 1.5
    case = dict(dim=2, dtype=dtype, nx=nx, ny=ny, gc=2, bc="periodic", pure=True, theta=1.2,
                cfl=0.25, stepper="fe", recon="linear")
    return case, (w, np.zeros_like(w), np.zeros_like(w))


def dry_states_2d(dtype="float64", n=64, stepper="rk2"):
    syn = synthetic()
    x = syn.cell_centers(n, 2, 0.0, 1.0)
    r2 = (x[None, :] - 0.5) ** 2 + (x[:, None] - 0.5) ** 2
    w = 1.0 - np.sign(r2 - 0.01)
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc="periodic", theta=1.2, cfl=0.25,
                stepper=stepper, recon="linear", depth_cutoff=1e-5, kappa=1e-5)
    return case, (w, np.zeros_like(w), np.zeros_like(w))


# ------------------------------------------------------------------------------------------
# test_rain.jl / test_infiltration.jl
# ------------------------------------------------------------------------------------------
def rain_case(n=10, rain=0.01, dtype="float64"):
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc="wall", extent=[0.0, 1000.0, 0.0, 1000.0],
                theta=1.2, cfl=0.25, stepper="fe", recon="linear", rain=rain)
    w = np.ones((n, n))
    return case, (w, np.zeros_like(w), np.zeros_like(w))


# test_rain.jl:99-101, 112-122, 133-140: the closures handed to FunctionalRain
def constant_rain_function(t, x, y):
    return 0.01 / 3600.0


def timedependent_rain_function(t, x, y):
    if t < 1 * 3600:
        rate = 15.0 * 0.001
    elif t < 2 * 3600:
        rate = 0.3 * 0.001
    else:
        rate = 20.0 * 0.001
    return rate / 3600


def spatial_rain_function(t, x, y):
    if t > 3600.0 and t < 7200.0:
        return 3 * (x / 500) * (y / 500) * 0.01 / 3600.0
    return 0.0


def infiltration_case(n=10, dtype="float64", t0=1e6):
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc="periodic", extent=[0.0, 100.0, 0.0, 100.0],
                theta=1.2, cfl=0.25, stepper="fe", recon="linear", infiltration={}, t0=t0)
    w = np.ones((n, n))
    return case, (w, np.zeros_like(w), np.zeros_like(w))


# ------------------------------------------------------------------------------------------
# 1D Gaussian (examples/shallow_water_1d.jl:31-66 = BASELINE config 1; test_swe1D_sim.jl)
# ------------------------------------------------------------------------------------------
def gaussian_1d(pure=True, B=None, elevate=0.0, stepper="fe", cfl=0.2, friction=None,
                theta=1.05, nx=1024, bottom_source=False, dtype="float64"):
    syn = synthetic()
    x = syn.cell_centers(nx, 2, 0.0, 1.0)
    h = np.exp(-((x - 0.5) ** 2) / 0.001) + 1.5 + elevate
    case = dict(dim=1, dtype=dtype, nx=nx, gc=2, bc="periodic", extent=[0.0, 1.0], pure=pure,
                theta=theta, cfl=cfl, stepper=stepper, recon="linear", friction=friction,
                B=B, bottom_source=bottom_source)
    return case, (h, np.zeros_like(h))


# ------------------------------------------------------------------------------------------
# BASELINE configs 3-5 on synthetic terrain
# ------------------------------------------------------------------------------------------
def terrain_case(n=4096, dtype="float32", bc="wall", friction="bsa2012", stepper="rk2", level=None,
                 wavelength=512.0, film=0.0):
    """config 3: terrain.jl-shaped, bottom source + Manning-type friction + rain + Horton."""
    syn = synthetic()
    npd = np.dtype(dtype)
    B = syn.terrain_corners(n, n, 2, seed=1234, dtype=npd, wavelength=wavelength)
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc=bc, extent=[0.0, float(n), 0.0, float(n)],
                depth_cutoff=8e-2, kappa=1e-5, theta=1.2, cfl=0.25, stepper=stepper, recon="linear",
                B=B, bottom_source=True, rain=0.015, infiltration={}, infil_first=True,
                friction=friction, maximum_timestep=1.0)
    if level is None:
        w = (syn.cell_average(B)[2:-2, 2:-2] + film).astype(npd)  # IC w = B_cell (dry) [+ a uniform film]
        state = (w, np.zeros_like(w), np.zeros_like(w))
    else:
        state = syn.lake_initial_state(B, 2, level, dtype=npd)
    return case, state


def urban_case(n=8192, dtype="float32", bc="wall", level=None, wavelength=512.0):
    """config 4: urban.jl-style, NoReconstruction, rain + constant infiltration, bsa2012 friction."""
    syn = synthetic()
    npd = np.dtype(dtype)
    B = syn.add_buildings(syn.terrain_corners(n, n, 2, seed=1234, dtype=npd, wavelength=wavelength), seed=4321)
    case = dict(dim=2, dtype=dtype, nx=n, ny=n, gc=2, bc=bc, extent=[0.0, float(n), 0.0, float(n)],
                depth_cutoff=8e-2, kappa=8e-2, cfl=0.1, stepper="rk2", recon="none",
                B=B, bottom_source=True, rain=15.0 / 1000.0, infiltration=15.0 / 1000.0 / 3600.0,
                infil_first=True, friction="bsa2012", maximum_timestep=1.0)
    if level is None:
        w = syn.cell_average(B)[2:-2, 2:-2].astype(npd)
        state = (w, np.zeros_like(w), np.zeros_like(w))
    else:
        state = syn.lake_initial_state(B, 2, level, dtype=npd)
    return case, state


def lake_case(n=16384, dtype="float32", level=0.0, wavelength=512.0, ny=None, cutoff=1e-3):
    """config 5: WallBC, minmod 1.2, bottom source, bh2021 friction, RK2, cfl .25, partially wet
    (~55 % of the cells).  depth_cutoff = kappa = 1e-3 (1 mm): at |w| ~ 10 m one Float32 ulp of the
    stored free surface is 1e-6 m, so the reference's Float64-oriented default 1e-5 would sit only
    10 ulps above the storage resolution."""
    syn = synthetic()
    npd = np.dtype(dtype)
    ny = n if ny is None else ny
    B = syn.terrain_corners(n, ny, 2, seed=1234, dtype=npd, wavelength=wavelength)
    case = dict(dim=2, dtype=dtype, nx=n, ny=ny, gc=2, bc="wall", extent=[0.0, float(n), 0.0, float(ny)],
                depth_cutoff=cutoff, kappa=cutoff, theta=1.2, cfl=0.25, stepper="rk2", recon="linear",
                B=B, bottom_source=True, friction="bh2021")
    state = syn.lake_initial_state(B, 2, level, dtype=npd)
    return case, state


def perturb_ulp(state, dtype, seed=7):
    """+-1 ulp (of the storage dtype) on a random half of the entries of each field: the smallest input
    change the storage type can express.  Used to measure how ill-conditioned N steps of a case are."""
    rng = np.random.default_rng(seed)
    out = []
    for a in state:
        a = np.ascontiguousarray(a, dtype=dtype).copy()
        up = np.nextafter(a, np.asarray(np.inf, dtype=dtype))
        dn = np.nextafter(a, np.asarray(-np.inf, dtype=dtype))
        r = rng.integers(0, 4, size=a.shape)
        a = np.where(r == 0, up, np.where(r == 1, dn, a))
        out.append(a.astype(dtype))
    # keep exact zeros exact (dry momentum): a subnormal momentum is not a meaningful perturbation
    return tuple(np.where(np.asarray(s0) == 0, 0, a).astype(dtype) for a, s0 in zip(out, state))


def field_errors(a, b):
    """rel L-inf per field: w against max|w|, both momenta against max(|hu|,|hv|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    dw = max(np.max(np.abs(b[0])), 1e-300)
    dm = max(np.max(np.abs(b[1])), np.max(np.abs(b[2])), 1e-300)
    return [float(np.max(np.abs(a[0] - b[0])) / dw), float(np.max(np.abs(a[1] - b[1])) / dm),
            float(np.max(np.abs(a[2] - b[2])) / dm)]


def lake_strip_case(n, rank, nranks, dtype="float32", level=0.0, wavelength=512.0, cutoff=1e-3, rows=None):
    """config 5 generated directly as the y-strip of `rank` (no rank ever materialises the global arrays).
    `rows` = (y_offset, n_rows) overrides the equal-height partition (load-balanced strips).
    Returns (strip case dict, strip state, y_offset)."""
    from_pkg = load_package()
    syn = from_pkg.synthetic
    if rows is None:
        offsets, sizes = from_pkg.multigpu.partition_rows(n, nranks)
        y0, nr = offsets[rank], sizes[rank]
    else:
        y0, nr = rows
    npd = np.dtype(dtype)
    B = syn.terrain_corners(n, n, 2, seed=1234, dtype=npd, wavelength=wavelength, row_range=(y0, y0 + nr + 5))
    case = dict(dim=2, dtype=dtype, nx=n, ny=nr, gc=2, bc="wall", extent=[0.0, float(n), 0.0, float(n)],
                depth_cutoff=cutoff, kappa=cutoff, theta=1.2, cfl=0.25, stepper="rk2", recon="linear",
                B=B, bottom_source=True, friction="bh2021")
    state = syn.lake_initial_state(B, 2, level, dtype=npd, y_offset=y0, ny_global=n)
    return case, state, y0


def rel_linf(a, b):
    """SURVEY 8c parity protocol: max|a-b| / max|b| per field."""
    den = np.max(np.abs(b))
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))) /
                 (den if den > 0 else 1.0))
