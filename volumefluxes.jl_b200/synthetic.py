"""Grid coordinate helpers and seeded synthetic inputs (terrain, initial conditions).

The reference's example terrain (SWIM artifact data, Artifacts.toml:1-6) is not
downloadable offline, so BASELINE.json's configs 3-5 run on synthetic terrain of the same
shape.  Coordinates follow grid.jl:92-161 (LinRange-based faces, centres = face + delta/2).
All arrays are returned in numpy C-order shaped (Ny, Nx): byte-identical to the
reference's column-major (Nx, Ny).
"""
from __future__ import annotations

import numpy as np


def linrange(start: float, stop: float, length: int) -> np.ndarray:
    """Julia LinRange(start, stop, length): lerp (1-t)*start + t*stop, t = j/(length-1)."""
    j = np.arange(length, dtype=np.float64)
    t = j / (length - 1)
    return (1.0 - t) * start + t * stop


def cell_faces(n: int, gc: int, lo: float, hi: float, interior: bool = True) -> np.ndarray:
    """grid.jl:92-115 -- face coordinates along one direction."""
    if interior:
        return linrange(lo, hi, n + 1)
    dx = (hi - lo) / n
    return linrange(lo - gc * dx, hi + gc * dx, n + 2 * gc + 1)


def cell_centers(n: int, gc: int, lo: float, hi: float, interior: bool = True) -> np.ndarray:
    """grid.jl:139-161 -- centres = faces[:-1] + delta/2 with delta = (hi-lo)/n."""
    f = cell_faces(n, gc, lo, hi, interior)
    return f[:-1] + ((hi - lo) / n) / 2.0


def terrain_corners(nx: int, ny: int, gc: int = 2, extent=None, seed: int = 1234,
                    amplitude: float = None, wavelength: float = 512.0, slope: float = 0.0,
                    depression: float = None, dtype=np.float32, octaves: int = 5, row_range=None) -> np.ndarray:
    """Smooth hills (sum of `octaves` separable sinusoid products with seeded phases) + a linear
    slope in x + one bay-like Gaussian depression.  Evaluated at the (nx+2gc+1) x (ny+2gc+1) cell
    corners (incl. ghost region), as BottomTopography2D expects (bottom_topography.jl:51-59).
    Every term is separable, so the field is one small GEMM: cheap even at 16389^2."""
    if extent is None:
        extent = [0.0, float(nx), 0.0, float(ny)]
    if amplitude is None:
        amplitude = 10.0 * wavelength / 512.0     # keeps hill slopes (~0.1 rms) independent of the wavelength
    if depression is None:
        depression = 0.375 * amplitude
    x = cell_faces(nx, gc, extent[0], extent[1], interior=False)
    y = cell_faces(ny, gc, extent[2], extent[3], interior=False)
    Lx, Ly = extent[1] - extent[0], extent[3] - extent[2]
    rng = np.random.default_rng(seed)
    ax, by = [], []
    norm = sum(0.5 ** o for o in range(octaves))
    for o in range(octaves):
        kx = 2 * np.pi * (2 ** o) / wavelength
        ph = rng.uniform(0, 2 * np.pi, size=4)
        rot = rng.uniform(0.6, 1.4)
        a = amplitude * (0.5 ** o) / norm
        # two products per octave so the pattern is not axis-symmetric
        ax.append(a * 0.5 * np.sin(kx * x + ph[0])); by.append(np.cos(kx * rot * y + ph[1]))
        ax.append(a * 0.5 * np.cos(kx * rot * x + ph[2])); by.append(np.sin(kx * y + ph[3]))
    ax.append(slope * (x - extent[0])); by.append(np.ones_like(y))
    s2x, s2y = (0.15 * Lx) ** 2, (0.15 * Ly) ** 2
    ax.append(-depression * np.exp(-((x - (extent[0] + 0.6 * Lx)) ** 2) / s2x))
    by.append(np.exp(-((y - (extent[2] + 0.4 * Ly)) ** 2) / s2y))
    A = np.stack(ax, axis=0).astype(dtype)      # (K, Nx+1)
    Bm = np.stack(by, axis=1).astype(dtype)     # (Ny+1, K)
    if row_range is not None:                   # only the corner rows [r0, r1) (a y-strip of the field)
        Bm = Bm[row_range[0]:row_range[1]]
    return np.ascontiguousarray(Bm @ A)          # (Ny+1, Nx+1)


def add_buildings(B: np.ndarray, seed: int = 4321, coverage: float = 0.15, size=(12, 40),
                  height=(6.0, 20.0)) -> np.ndarray:
    """Axis-aligned rectangular plateaus (urban.jl-style DSM buildings) on a corner field."""
    rng = np.random.default_rng(seed)
    out = B.copy()
    ny, nx = B.shape
    mean_area = ((size[0] + size[1]) / 2.0) ** 2
    n = int(coverage * nx * ny / mean_area)
    xs = rng.integers(0, nx, n); ys = rng.integers(0, ny, n)
    ws = rng.integers(size[0], size[1], n); hs = rng.integers(size[0], size[1], n)
    zs = rng.uniform(height[0], height[1], n)
    for x0, y0, w, h, z in zip(xs, ys, ws, hs, zs):
        out[y0:y0 + h, x0:x0 + w] = B[y0:y0 + h, x0:x0 + w] + out.dtype.type(z)
    return out


def cell_average(Bc: np.ndarray) -> np.ndarray:
    """B_cell (bottom_topography.jl:97): 0.25*(sum of the 4 corners), same summation order."""
    s = Bc[:-1, :-1] + Bc[:-1, 1:]       # B[i,j] + B[i+1,j]
    s = s + Bc[1:, :-1]                  # + B[i,j+1]
    s = s + Bc[1:, 1:]                   # + B[i+1,j+1]
    return 0.25 * s.astype(np.float64)


def lake_initial_state(Bc: np.ndarray, gc: int, level: float, hump: float = 0.5, dtype=np.float32,
                       y_offset: int = 0, ny_global: int = None):
    """w = max(B_cell, level) plus a Gaussian hump on the wet part; hu = hv = 0.  Interior only.
    With y_offset / ny_global, Bc is the corner strip of the rows [y_offset, y_offset + ny) of a taller grid."""
    Bcell = cell_average(Bc)[gc:-gc, gc:-gc]
    ny, nx = Bcell.shape
    y = (y_offset + np.arange(ny) + 0.5) / (ny_global if ny_global is not None else ny)
    x = (np.arange(nx) + 0.5) / nx
    bump = hump * np.exp(-((x[None, :] - 0.6) ** 2 + (y[:, None] - 0.4) ** 2) / 0.01)
    w = np.maximum(Bcell, level + bump)
    w = w.astype(dtype)
    z = np.zeros_like(w)
    return w, z, z.copy()
