"""The workloads of BASELINE.json's configs as plain case dicts + initial states (synthetic data: the reference's
terrain artifacts are not downloadable offline).  Used by bench.py, the tests (tests/cases.py re-exports them) and
scratch/parity_report.py.  A case dict is what `api.simulator_from_case` and the CPU oracle both understand.

Arrays are numpy C-order (Ny, Nx) == the reference's column-major (Nx, Ny).

Deviation from SURVEY.md section 8(d), config 5: the survey sketched `kappa = depth_cutoff = 1e-5` and a lake level of
5 m on a 40 m terrain; the generator here has 5-octave hills of about +-8 m, so the lake level is 0 (54 % of the cells
wet) and `depth_cutoff = kappa = 1e-3`: at |w| ~ 10 m one Float32 ulp of the stored free surface is 1e-6 m, so the
reference's Float64-oriented default 1e-5 would sit only 10 ulps above the storage resolution (bench.py states this in
`config.workload`)."""
from __future__ import annotations

import numpy as np

from . import multigpu, synthetic


def swashes41x(kind="412", nx=512, layout="x", dtype="float64", stepper="fe", ny=3):
    L, x0, hl = 10.0, 5.0, 0.005
    hr = 0.0 if kind == "412" else 0.001
    syn = synthetic
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

def terrain_case(n=4096, dtype="float32", bc="wall", friction="bsa2012", stepper="rk2", level=None,
                 wavelength=512.0, film=0.0):
    """config 3: terrain.jl-shaped, bottom source + Manning-type friction + rain + Horton."""
    syn = synthetic
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
    syn = synthetic
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
    syn = synthetic
    npd = np.dtype(dtype)
    ny = n if ny is None else ny
    B = syn.terrain_corners(n, ny, 2, seed=1234, dtype=npd, wavelength=wavelength)
    case = dict(dim=2, dtype=dtype, nx=n, ny=ny, gc=2, bc="wall", extent=[0.0, float(n), 0.0, float(ny)],
                depth_cutoff=cutoff, kappa=cutoff, theta=1.2, cfl=0.25, stepper="rk2", recon="linear",
                B=B, bottom_source=True, friction="bh2021")
    state = syn.lake_initial_state(B, 2, level, dtype=npd)
    return case, state

def lake_strip_case(n, rank, nranks, dtype="float32", level=0.0, wavelength=512.0, cutoff=1e-3, rows=None):
    """config 5 generated directly as the y-strip of `rank` (no rank ever materialises the global arrays).
    `rows` = (y_offset, n_rows) overrides the equal-height partition (load-balanced strips).
    Returns (strip case dict, strip state, y_offset)."""
    syn = synthetic
    if rows is None:
        offsets, sizes = multigpu.partition_rows(n, nranks)
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
