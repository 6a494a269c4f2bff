"""ctypes wrapper around oracle/liborc.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  The product package (volumefluxes.jl_b200) never does.

The oracle restates the reference's CPU algorithm (see swe_oracle.cpp header); this file
only marshals numpy arrays.  Arrays are column-major like the reference
(`data[i, j, v]`, x fastest): numpy arrays here are shaped (nvar, Ny, Nx) C-order, which
is byte-identical to Julia's (Nx, Ny, nvar).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


class OrcConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dtype", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("gc", C.c_int32),
        ("bc", C.c_int32), ("eq_kind", C.c_int32), ("recon", C.c_int32),
        ("stepper", C.c_int32), ("friction", C.c_int32),
        ("has_bathymetry", C.c_int32), ("has_bottom_source", C.c_int32),
        ("rain_kind", C.c_int32), ("infil_kind", C.c_int32), ("infil_first", C.c_int32),
        ("n_rain", C.c_int32),
        ("extent", C.c_double * 4),
        ("theta", C.c_double), ("rho", C.c_double), ("g", C.c_double),
        ("depth_cutoff", C.c_double), ("kappa", C.c_double), ("cfl", C.c_double),
        ("Cz", C.c_double), ("const_B", C.c_double), ("t0", C.c_double),
        ("rain_rate", C.c_double), ("infil_const", C.c_double),
        ("f0", C.c_double), ("fc", C.c_double), ("k", C.c_double),
        ("rain_rates", C.c_double * 16), ("rain_times", C.c_double * 16),
    ]


def build(force: bool = False) -> Path:
    so = _HERE / "liborc.so"
    src = _HERE / "swe_oracle.cpp"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liborc.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_bathymetry.argtypes = [vp, vp]
        L.orc_set_infiltration_factor.argtypes = [vp, vp]
        L.orc_set_rain_plane.argtypes = [vp, vp]
        L.orc_set_state_interior.argtypes = [vp, vp, vp, vp]
        L.orc_set_state_full.argtypes = [vp, i32, vp]
        L.orc_get_state_full.argtypes = [vp, i32, vp]
        L.orc_get_faces.argtypes = [vp, vp, vp]
        L.orc_get_wavespeeds.argtypes = [vp, vp]
        L.orc_get_time.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
        L.orc_set_time.argtypes = [vp, dbl, dbl]
        L.orc_step.argtypes = [vp, dbl]
        L.orc_step_advance.argtypes = [vp, dbl]
        L.orc_simulate_to_time.restype = i64
        L.orc_simulate_to_time.argtypes = [vp, dbl, i32, dbl, i64]
        L.orc_update_bc.argtypes = [vp, i32]
        L.orc_reconstruct.argtypes = [vp, i32, i32]
        L.orc_add_time_derivative.argtypes = [vp, i32, i32, dbl, C.POINTER(dbl)]
        L.orc_implicit_friction.argtypes = [vp, i32, i32, dbl]
        L.orc_post_proc.argtypes = [vp, i32]
        L.orc_minmod.restype = dbl
        L.orc_minmod.argtypes = [dbl, dbl, dbl]
        for name in ("orc_B_cell",):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [vp, i64, i64]
        for name in ("orc_B_face_left", "orc_B_face_right"):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [vp, i64, i64, i32]
        L.orc_desingularize.restype = dbl
        L.orc_desingularize.argtypes = [vp, dbl]
        L.orc_rain_rate.restype = dbl
        L.orc_rain_rate.argtypes = [vp, dbl]
        L.orc_infiltration.restype = dbl
        L.orc_infiltration.argtypes = [vp, dbl, i64, i64]
        L.orc_enumerate_inner_cells.restype = i64
        L.orc_enumerate_inner_cells.argtypes = [vp, i32, i32, vp]
        L.orc_enumerate_ghost_cells.restype = i64
        L.orc_enumerate_ghost_cells.argtypes = [vp, i32, vp]
        L.orc_enumerate_cells.restype = i64
        L.orc_enumerate_cells.argtypes = [vp, vp]
        L.orc_zero.argtypes = [vp, i32]
        L.orc_axpy.argtypes = [vp, i32, i32, dbl, i32]
        L.orc_rk_average.argtypes = [vp, i32, i32]
        L.orc_update_bc_sides.argtypes = [vp, i32, i32, i32, i32, i32]
        L.orc_rotate.argtypes = [vp]
        L.orc_set_threads.restype = i32
        L.orc_set_threads.argtypes = [i32]
        _lib = L
    return _lib


BC = {"periodic": 0, "wall": 1, "neumann": 2}
FRICTION = {None: 0, "none": 0, "bh2021": 1, "bsa2012": 2, "fcg2016": 3}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """One reference-style Simulator on the CPU oracle.  `case` is a plain dict (see tests/cases.py)."""

    def __init__(self, case: dict):
        c = OrcConfig()
        self.dim = c.dim = int(case.get("dim", 2))
        self.np_dtype = np.dtype(case.get("dtype", "float64"))
        c.dtype = 0 if self.np_dtype == np.float32 else 1
        c.nx = int(case["nx"])
        c.ny = int(case.get("ny", 1))
        c.gc = int(case.get("gc", 2))
        c.bc = BC[case.get("bc", "periodic")]
        c.eq_kind = 1 if case.get("pure", False) else 0
        c.recon = 1 if case.get("recon", "linear") == "linear" else 0
        c.stepper = 1 if case.get("stepper", "rk2") == "rk2" else 0
        c.friction = FRICTION[case.get("friction")]
        B = case.get("B")
        c.has_bathymetry = 0 if B is None or np.isscalar(B) else 1
        c.const_B = float(B) if (B is not None and np.isscalar(B)) else 0.0
        c.has_bottom_source = 1 if case.get("bottom_source", False) else 0
        rain = case.get("rain")
        if rain is None:
            c.rain_kind = 0
        elif isinstance(rain, dict):      # FunctionalRain(rain_function, grid), source_term_rain.jl:57-68
            c.rain_kind = 3
        elif np.isscalar(rain):
            c.rain_kind, c.rain_rate = 1, float(rain)
        else:
            rates, times = rain
            c.rain_kind, c.n_rain = 2, len(rates)
            for i, (r, t) in enumerate(zip(rates, times)):
                c.rain_rates[i], c.rain_times[i] = float(r), float(t)
        infil = case.get("infiltration")
        if infil is None:
            c.infil_kind = 0
        elif np.isscalar(infil):
            c.infil_kind, c.infil_const = 1, float(infil)
        else:
            c.infil_kind = 2
            c.f0, c.fc, c.k = (float(infil.get("f0", 1.977e-4)), float(infil.get("fc", 3.272e-5)),
                               float(infil.get("k", 2.43e-3)))
        c.infil_first = 1 if case.get("infil_first", True) else 0
        ext = case.get("extent", [0.0, 1.0, 0.0, 1.0] if self.dim == 2 else [0.0, 1.0, 0.0, 0.0])
        for i in range(4):
            c.extent[i] = float(ext[i]) if i < len(ext) else 0.0
        c.theta = float(case.get("theta", 1.2))
        c.rho = float(case.get("rho", 1.0))
        c.g = float(case.get("g", 9.81))
        c.depth_cutoff = float(case.get("depth_cutoff", 1e-5))
        c.kappa = float(case.get("kappa", 1e-5))
        c.cfl = float(case.get("cfl", 0.25))
        c.Cz = float(case.get("Cz", 0.03 ** 2))
        c.t0 = float(case.get("t0", 0.0))
        self.cfg = c
        self.nvar = 2 if self.dim == 1 else 3
        self.Nx = c.nx + 2 * c.gc
        self.Ny = 1 if self.dim == 1 else c.ny + 2 * c.gc
        self._h = lib().orc_create(C.byref(c))
        if c.has_bathymetry:
            Ba = np.ascontiguousarray(B, dtype=self.np_dtype)
            want = (self.Nx + 1,) if self.dim == 1 else (self.Ny + 1, self.Nx + 1)
            if Ba.shape != want:
                raise ValueError(f"Bottom topography should be of shape {want} but got {Ba.shape}")
            lib().orc_set_bathymetry(self._h, _ptr(Ba))
        if c.infil_kind == 2 and not np.isscalar(infil.get("factor", 1.0)):
            f = np.ascontiguousarray(infil["factor"], dtype=np.float64)
            assert f.shape == (self.Ny, self.Nx)
            lib().orc_set_infiltration_factor(self._h, _ptr(f))

        self._rain_fn = rain["function"] if isinstance(rain, dict) else None
        if self._rain_fn is not None:
            assert self.dim == 2
            self._rain_t = None
            self._refresh_rain(c.t0)

    def _refresh_rain(self, t):
        """compute_rain(rain::FunctionalRain, t, index) = rain_function(t, cell_center(grid, index)...)
        (source_term_rain.jl:65-68, grid.jl:117-121), evaluated for every cell like the reference's for_each_cell."""
        if self._rain_fn is None or t == self._rain_t:
            return
        c = self.cfg
        dx = (c.extent[1] - c.extent[0]) / c.nx
        dy = (c.extent[3] - c.extent[2]) / c.ny
        v = np.empty((self.Ny, self.Nx), dtype=np.float64)
        for j in range(self.Ny):
            y = c.extent[2] + dy * ((j + 1) - 0.5 - c.gc)
            for i in range(self.Nx):
                x = c.extent[0] + dx * ((i + 1) - 0.5 - c.gc)
                v[j, i] = self._rain_fn(t, x, y)
        lib().orc_set_rain_plane(self._h, _ptr(v))
        self._rain_t = t

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- state ---------------------------------------------------------------------------
    def set_state_interior(self, w, hu, hv=None):
        arrs = [np.ascontiguousarray(a, dtype=self.np_dtype) for a in (w, hu)]
        arrs.append(np.ascontiguousarray(hv if hv is not None else np.zeros_like(arrs[0]), dtype=self.np_dtype))
        lib().orc_set_state_interior(self._h, *[_ptr(a) for a in arrs])

    def full_shape(self):
        return (self.nvar, self.Nx) if self.dim == 1 else (self.nvar, self.Ny, self.Nx)

    def get_state_full(self, which=0):
        out = np.empty(self.full_shape(), dtype=self.np_dtype)
        lib().orc_get_state_full(self._h, which, _ptr(out))
        return out

    def set_state_full(self, data, which=0):
        d = np.ascontiguousarray(data, dtype=self.np_dtype)
        assert d.shape == self.full_shape()
        lib().orc_set_state_full(self._h, which, _ptr(d))

    def get_state_interior(self):
        f = self.get_state_full()
        g = self.cfg.gc
        if self.dim == 1:
            return f[:, g:-g].copy()
        return f[:, g:-g, g:-g].copy()

    def get_faces(self):
        l = np.empty(self.full_shape(), dtype=self.np_dtype)
        r = np.empty(self.full_shape(), dtype=self.np_dtype)
        lib().orc_get_faces(self._h, _ptr(l), _ptr(r))
        return l, r

    def get_wavespeeds(self):
        w = np.empty(self.full_shape()[1:], dtype=self.np_dtype)
        lib().orc_get_wavespeeds(self._h, _ptr(w))
        return w

    # -- time ----------------------------------------------------------------------------
    @property
    def t(self):
        t, dt = C.c_double(), C.c_double()
        lib().orc_get_time(self._h, C.byref(t), C.byref(dt))
        return t.value

    @property
    def dt(self):
        t, dt = C.c_double(), C.c_double()
        lib().orc_get_time(self._h, C.byref(t), C.byref(dt))
        return dt.value

    def set_time(self, t, dt=0.0):
        lib().orc_set_time(self._h, float(t), float(dt))

    def step(self, max_dt=float("inf"), advance=True):
        self._refresh_rain(self.t)
        (lib().orc_step_advance if advance else lib().orc_step)(self._h, float(max_dt))

    def simulate_to_time(self, T, match_endtime=True, maximum_timestep=None, max_steps=0):
        if self._rain_fn is not None:      # simulator.jl:179-212 as a host loop: the closure is re-evaluated every step
            n = 0
            while self.t < T:
                max_dt = (T - self.t) if match_endtime else float("inf")
                if maximum_timestep:
                    max_dt = min(max_dt, maximum_timestep)
                self.step(max_dt)
                n += 1
            return n
        return lib().orc_simulate_to_time(self._h, float(T), int(match_endtime),
                                          float(maximum_timestep or 0.0), int(max_steps))

    # -- fine-grained operators ----------------------------------------------------------------
    def update_bc(self, which=0):
        lib().orc_update_bc(self._h, which)

    def reconstruct(self, direction, which=0):
        lib().orc_reconstruct(self._h, direction, which)

    def add_time_derivative(self, in_buffer=0, out_buffer=1, t=0.0):
        ms = (C.c_double * 2)()
        lib().orc_add_time_derivative(self._h, in_buffer, out_buffer, float(t), ms)
        return ms[0], ms[1]

    def implicit_friction(self, out_buffer, prev_buffer, dt):
        lib().orc_implicit_friction(self._h, out_buffer, prev_buffer, float(dt))

    def post_proc(self, which):
        lib().orc_post_proc(self._h, which)

    # pieces of do_substep! / perform_step! for host-driven (y-strip) runs
    def zero(self, which):
        lib().orc_zero(self._h, which)

    def axpy(self, out_buffer, state_buffer, dt, rounded=False):
        lib().orc_axpy(self._h, out_buffer, state_buffer, float(dt), 1 if rounded else 0)

    def rk_average(self, out_buffer, first_buffer=0):
        lib().orc_rk_average(self._h, out_buffer, first_buffer)

    def update_bc_sides(self, which, xlo=True, xhi=True, ylo=True, yhi=True):
        lib().orc_update_bc_sides(self._h, which, int(xlo), int(xhi), int(ylo), int(yhi))

    def rotate(self):
        lib().orc_rotate(self._h)

    # -- scalars / index spaces ---------------------------------------------------------------
    def B_cell(self, i, j=1):
        return lib().orc_B_cell(self._h, i - 1, j - 1)

    def B_face_left(self, i, j=1, direction=0):
        return lib().orc_B_face_left(self._h, i - 1, j - 1, direction)

    def B_face_right(self, i, j=1, direction=0):
        return lib().orc_B_face_right(self._h, i - 1, j - 1, direction)

    def desingularize(self, h):
        return lib().orc_desingularize(self._h, float(h))

    def rain_rate(self, t):
        return lib().orc_rain_rate(self._h, float(t))

    def infiltration(self, t, i=1, j=1):
        return lib().orc_infiltration(self._h, float(t), i - 1, j - 1)

    def enumerate_inner_cells(self, direction, ghostcells=None):
        g = self.cfg.gc if ghostcells is None else ghostcells
        n = lib().orc_enumerate_inner_cells(self._h, direction, g, None)
        out = np.empty((n, 6), dtype=np.int64)
        lib().orc_enumerate_inner_cells(self._h, direction, g, _ptr(out))
        return out

    def enumerate_ghost_cells(self, direction):
        n = lib().orc_enumerate_ghost_cells(self._h, direction, None)
        out = np.empty((n, 2), dtype=np.int64)
        lib().orc_enumerate_ghost_cells(self._h, direction, _ptr(out))
        return out

    def enumerate_cells(self):
        n = lib().orc_enumerate_cells(self._h, None)
        out = np.empty((n, 2), dtype=np.int64)
        lib().orc_enumerate_cells(self._h, _ptr(out))
        return out


def set_threads(n: int) -> int:
    """host threads used by the per-cell loops (0 = leave unchanged); returns the count in effect"""
    return lib().orc_set_threads(int(n))


def minmod(a, b, c):
    return lib().orc_minmod(float(a), float(b), float(c))
