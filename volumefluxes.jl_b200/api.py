"""Host-side mirror of the reference's operator interface for the 2D shallow-water step, over the
C ABI (include/swe_b200.h).  Names, argument meaning and error behaviour follow SinFVM
(paths below are relative to the reference's src/):

    CartesianGrid, PeriodicBC/WallBC/NeumannBC                      grid.jl:21-90
    make_b200_backend (next to make_cpu_backend/make_cuda_backend)  backends/kernel_abstractions.jl:29-55
    for_each_cell / for_each_inner_cell / for_each_ghost_cell       backends/kernel_abstractions.jl:66-131
    ConstantBottomTopography, BottomTopography2D                    bottom_topography.jl:24-107
    ShallowWaterEquations, ShallowWaterEquationsPure                equations/swe_2D.jl, swe_2D_pure.jl
    LinearReconstruction, NoReconstruction                          reconstruction/*.jl
    CentralUpwind                                                   numericalflux/swe/centralupwind.jl
    SourceTermBottom, ConstantRain, TimeDependentRain,
    HortonInfiltration, ConstantInfiltration                        sourceterms/*.jl
    ImplicitFriction, friction_bh2021/bsa2012/fcg2016               friction.jl
    ConservedSystem                                                 system.jl
    RungeKutta2, ForwardEulerStepper                                timestepper/*.jl
    Simulator, set_current_state, current_interior_state,
    perform_step, simulate_to_time, update_bc                       simulator.jl, bc.jl

Array convention: numpy arrays are C-order with shape (ny, nx) -- byte-identical to the
reference's column-major (nx, ny) -- so `a[j, i]` here is `a[i+1, j+1]` in Julia.

Every compute call goes to the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from . import synthetic as _syn

__all__ = [
    "XDIR", "YDIR", "PeriodicBC", "WallBC", "NeumannBC", "CartesianGrid", "B200Backend",
    "make_b200_backend", "get_available_backends", "name", "DomainError", "ArgumentError", "B200Error",
    "ConstantBottomTopography", "BottomTopography2D", "is_zero", "ShallowWaterEquations",
    "ShallowWaterEquationsPure", "LinearReconstruction", "NoReconstruction", "CentralUpwind",
    "SourceTermBottom", "ConstantRain", "TimeDependentRain", "FunctionalRain", "HortonInfiltration",
    "ConstantInfiltration", "ImplicitFriction", "friction_bh2021", "friction_bsa2012", "friction_fcg2016",
    "ConservedSystem", "RungeKutta2", "ForwardEulerStepper", "Simulator", "set_current_state",
    "current_interior_state", "current_state", "current_time", "current_timestep", "perform_step",
    "simulate_to_time", "update_bc", "for_each_cell", "for_each_inner_cell", "for_each_ghost_cell",
    "cell_faces", "cell_centers", "compute_dx", "compute_dy", "inner_cells", "interior_size",
    "simulator_from_case", "IntervalWriter", "TotalWaterVolume", "AsyncIntervalWriter", "Snapshot", "pinned_empty", "BottomTopography1D",
    "ShallowWaterEquations1D", "ShallowWaterEquations1DPure",
]

XDIR, YDIR = 0, 1


class B200Error(RuntimeError):
    """CUDA / library failure (SWE_B200_ECUDA, SWE_B200_EUNSUPPORTED)."""


class DomainError(ValueError):
    """Julia's DomainError (bad bathymetry / infiltration-factor size)."""


class ArgumentError(ValueError):
    """Julia's ArgumentError / failed @assert."""


def _raise(rc: int, handle=None):
    msg = _lib.lib().swe_b200_last_error(handle).decode() if True else ""
    if rc == _lib.EDOMAIN:
        raise DomainError(msg)
    if rc == _lib.EINVAL:
        raise ArgumentError(msg)
    raise B200Error(f"swe_b200 error {rc}: {msg}")


# ---- grid.jl ------------------------------------------------------------------------------------
class PeriodicBC:
    code = 0


class WallBC:
    code = 1


class NeumannBC:
    code = 2


class CartesianGrid:
    """CartesianGrid(nx, ny; gc=1, boundary=PeriodicBC(), extent=[0 1; 0 1]) (grid.jl:82-90) or the 1D
    CartesianGrid(nx; gc=1, boundary=PeriodicBC(), extent=[0 1]) (grid.jl:72-80)."""

    def __init__(self, nx: int, ny: Optional[int] = None, gc: int = 1, boundary=None, extent=None):
        self.boundary = boundary if boundary is not None else PeriodicBC()
        self.dim = 1 if ny is None else 2
        if ny is None:
            ext = np.asarray(extent if extent is not None else [0.0, 1.0], dtype=np.float64).reshape(2)
            self.extent = (float(ext[0]), float(ext[1]))
            self.ghostcells = (int(gc),)
            self.totalcells = (int(nx) + 2 * gc,)
            self.delta = ((ext[1] - ext[0]) / nx,)
            self.nx, self.ny, self.gc = int(nx), None, int(gc)
            return
        ext = np.asarray(extent if extent is not None else [[0.0, 1.0], [0.0, 1.0]], dtype=np.float64).reshape(2, 2)
        self.extent = (float(ext[0, 0]), float(ext[0, 1]), float(ext[1, 0]), float(ext[1, 1]))
        self.ghostcells = (int(gc), int(gc))
        self.totalcells = (int(nx) + 2 * gc, int(ny) + 2 * gc)
        self.delta = ((ext[0, 1] - ext[0, 0]) / nx, (ext[1, 1] - ext[1, 0]) / ny)
        self.nx, self.ny, self.gc = int(nx), int(ny), int(gc)

    def size(self):
        return self.totalcells


def interior_size(grid):
    return (grid.nx, grid.ny)


def compute_dx(grid, direction=XDIR):
    return grid.delta[direction]


def compute_dy(grid):
    return grid.delta[YDIR]


def inner_cells(grid, direction, ghostcells=None):
    """grid.jl:213-220"""
    g = grid.ghostcells[direction] if ghostcells is None else ghostcells
    if direction == XDIR:
        return (grid.totalcells[0] - 2 * g, grid.totalcells[1] - 2 * grid.gc)
    return (grid.totalcells[0] - 2 * grid.gc, grid.totalcells[1] - 2 * g)


def cell_faces(grid, direction=None, interior=True):
    """grid.jl:92-135.  With a direction: 1-D face coordinates; without: (x, y) of each corner,
    shaped (ny+1, nx+1, 2) (or (Ny+1, Nx+1, 2) with interior=False)."""
    if direction is not None:
        n = grid.nx if direction == XDIR else grid.ny
        return _syn.cell_faces(n, grid.gc, grid.extent[2 * direction], grid.extent[2 * direction + 1], interior)
    x = cell_faces(grid, XDIR, interior)
    y = cell_faces(grid, YDIR, interior)
    return np.stack(np.meshgrid(x, y), axis=-1)


def cell_centers(grid, direction=None, interior=True):
    """grid.jl:139-161"""
    if direction is not None:
        n = grid.nx if direction == XDIR else grid.ny
        return _syn.cell_centers(n, grid.gc, grid.extent[2 * direction], grid.extent[2 * direction + 1], interior)
    x = cell_centers(grid, XDIR, interior)
    y = cell_centers(grid, YDIR, interior)
    return np.stack(np.meshgrid(x, y), axis=-1)


# ---- backends/kernel_abstractions.jl ---------------------------------------------------------------
@dataclass
class B200Backend:
    """A new `Backend` next to KernelAbstractionBackend (kernel_abstractions.jl:24-34): carries
    `realtype` like the reference's backends (read at system.jl:22, simulator.jl:80, buffer.jl:30)."""
    realtype: type = np.float64
    strict: bool = False      # bit-faithful arithmetic (see include/swe_b200.h)
    fused: bool = True        # fused stage kernels for perform_step!
    device: int = 0

    def __post_init__(self):
        self.realtype = np.dtype(self.realtype).type
        if self.realtype not in (np.float32, np.float64):
            raise ArgumentError("the B200 backend supports Float32 and Float64 (no Dual numbers)")
        _lib.lib()   # fail loudly right here if the CUDA library is missing


def make_b200_backend(realtype=np.float64, strict: bool = False, fused: bool = True, device: int = 0) -> B200Backend:
    return B200Backend(realtype, strict, fused, device)


def name(backend) -> str:
    return "B200"


def get_available_backends():
    """kernel_abstractions.jl:45-55 -- only the B200 backend lives in this package."""
    return [make_b200_backend()]


# ---- bottom_topography.jl ---------------------------------------------------------------------------
class ConstantBottomTopography:
    def __init__(self, B: float = 0.0):
        self.B = float(B)


class BottomTopography2D:
    """Corner values, shape (Ny+1, Nx+1) here == (Nx+1, Ny+1) column-major (bottom_topography.jl:51-59)."""

    def __init__(self, B, backend, grid):
        B = np.asarray(B)
        want = (grid.totalcells[1] + 1, grid.totalcells[0] + 1)
        if B.shape != want:   # validate_bottom_topography :78-82
            raise DomainError(f"Bottom topography should be of size {want[::-1]} but got {B.shape[::-1]}")
        self.B = np.ascontiguousarray(B, dtype=backend.realtype)


class BottomTopography1D:
    """Face values, shape (Nx+1,) (bottom_topography.jl:29-37)."""

    def __init__(self, B, backend, grid):
        B = np.asarray(B)
        want = (grid.totalcells[0] + 1,)
        if B.shape != want:   # validate_bottom_topography :78-82
            raise DomainError(f"Bottom topography should be of size {want} but got {B.shape}")
        self.B = np.ascontiguousarray(B, dtype=backend.realtype)


def is_zero(B) -> bool:
    return bool(np.all(np.asarray(B.B) == 0.0))


# ---- equations --------------------------------------------------------------------------------------
class ShallowWaterEquations:
    """equations/swe_2D.jl:22-29"""
    pure = False

    def __init__(self, B=None, rho=1.0, g=9.81, depth_cutoff=1e-5, desingularizing_kappa=1e-5):
        self.B = B if B is not None else ConstantBottomTopography()
        self.rho, self.g = float(rho), float(g)
        self.depth_cutoff, self.desingularizing_kappa = float(depth_cutoff), float(desingularizing_kappa)


class ShallowWaterEquationsPure:
    """equations/swe_2D_pure.jl:22-27"""
    pure = True

    def __init__(self, rho=1.0, g=9.81):
        self.rho, self.g = float(rho), float(g)
        self.B = ConstantBottomTopography()
        self.depth_cutoff = self.desingularizing_kappa = 1e-5


class ShallowWaterEquations1D(ShallowWaterEquations):
    """equations/swe_1D.jl:22-29 -- on this backend a 1D problem runs as a strip of identical rows of the 2D
    kernels (the reference asserts 1D run == x-run of the 2D solver bitwise, test_swashes.jl:317-328)."""


class ShallowWaterEquations1DPure(ShallowWaterEquationsPure):
    """equations/swe_1D_pure.jl:22-26"""


# 1D problems on the 2D kernels: ROWS_1D identical rows, y extent so large that dy / max|lambda_y| never limits dt
ROWS_1D = 2
DY_1D = 1.0e30


def _embed_1d(backend, system, grid):
    """(grid, system) of a 1D problem -> the equivalent strip of ROWS_1D identical rows: the y fluxes through the
    faces of identical rows cancel exactly, hv stays 0, and dt = cfl * min(dx/ax, dy/ay) = cfl * dx/ax."""
    g2 = CartesianGrid(grid.nx, ROWS_1D, gc=grid.gc, boundary=grid.boundary,
                       extent=[[grid.extent[0], grid.extent[1]], [0.0, ROWS_1D * DY_1D]])
    eq = system.equation
    if eq.pure:
        eq2 = ShallowWaterEquationsPure(eq.rho, eq.g)
    else:
        B = eq.B
        if isinstance(B, BottomTopography1D):
            B = BottomTopography2D(np.tile(B.B[None, :], (g2.totalcells[1] + 1, 1)), backend, g2)
        eq2 = ShallowWaterEquations(B, rho=eq.rho, g=eq.g, depth_cutoff=eq.depth_cutoff,
                                    desingularizing_kappa=eq.desingularizing_kappa)
    s2 = ConservedSystem(backend, system.reconstruction, CentralUpwind(eq2), eq2, g2, system.source_terms,
                         system.implicit_source_term)
    return s2, g2


class LinearReconstruction:
    def __init__(self, theta: float = 1.2):
        self.theta = float(theta)


class NoReconstruction:
    pass


class CentralUpwind:
    def __init__(self, eq):
        self.eq = eq


# ---- source terms / friction ------------------------------------------------------------------------
class SourceTermBottom:
    pass


class ConstantRain:
    def __init__(self, rain_rate):
        self.rain_rate = float(rain_rate)


class TimeDependentRain:
    def __init__(self, rain_rates, time=(0.0,)):
        if np.shape(rain_rates) != np.shape(time):   # source_term_rain.jl:36-38
            raise DomainError("Dimensions of input parameters rain_rates and time does not match")
        self.rain_rates = [float(r) for r in rain_rates]
        self.time = [float(t) for t in time]


class FunctionalRain:
    """FunctionalRain(rain_function, grid) (source_term_rain.jl:57-68): out.h += rain_function(t, x, y) at every cell
    centre.  A precompiled kernel cannot run the closure, so the host evaluates it over all cell centres
    (vectorised when the function accepts arrays) and uploads the values (swe_b200_set_rain_plane): once if
    `time_dependent=False`, before every step otherwise -- the simulation then runs the per-step host loop."""

    def __init__(self, rain_function: Callable, grid=None, time_dependent: bool = True):
        self.rain_function, self.grid, self.time_dependent = rain_function, grid, bool(time_dependent)

    def values(self, t: float, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        shape = (y.size, x.size)
        try:
            v = np.asarray(self.rain_function(t, x[None, :], y[:, None]), dtype=np.float64)
            v = np.broadcast_to(v, shape)
        except Exception:
            v = np.array([[self.rain_function(t, float(a), float(b)) for a in x] for b in y], dtype=np.float64)
        return np.ascontiguousarray(v)


class HortonInfiltration:
    """source_term_infiltration.jl:21-46; factor is a scalar or a (Ny, Nx) array over ALL cells."""

    def __init__(self, grid, backend, factor=1.0, fc=3.272e-5, f0=1.977e-4, k=2.43e-3):
        self.fc, self.f0, self.k = float(fc), float(f0), float(k)
        if np.ndim(factor) == 0:
            self.factor = np.full((grid.totalcells[1], grid.totalcells[0]), float(factor))
        else:
            f = np.asarray(factor, dtype=np.float64)
            want = (grid.totalcells[1], grid.totalcells[0])
            if f.shape != want:   # validate_infiltration_factor :61-65
                raise DomainError(f"Infiltration factor should be of size {want[::-1]} but got {f.shape[::-1]}")
            self.factor = np.ascontiguousarray(f)


class ConstantInfiltration:
    def __init__(self, infiltration_rate):
        self.infiltration_rate = float(infiltration_rate)


friction_bh2021, friction_bsa2012, friction_fcg2016 = "bh2021", "bsa2012", "fcg2016"
_FRICTION = {friction_bh2021: 1, friction_bsa2012: 2, friction_fcg2016: 3}


class ImplicitFriction:
    def __init__(self, Cz=0.03 ** 2, friction_function=friction_bh2021):
        if friction_function not in _FRICTION:
            raise B200Error("friction_function must be friction_bh2021 / friction_bsa2012 / friction_fcg2016")
        self.Cz, self.friction_function = float(Cz), friction_function


class RungeKutta2:
    code = 1


class ForwardEulerStepper:
    code = 0


# ---- system.jl ----------------------------------------------------------------------------------------
class ConservedSystem:
    def __init__(self, backend, reconstruction, numericalflux, equation, grid, source_terms=(),
                 implicit_source_term=None):
        if not isinstance(source_terms, (list, tuple)):
            source_terms = [source_terms]
        self.backend, self.reconstruction, self.numericalflux = backend, reconstruction, numericalflux
        self.equation, self.grid = equation, grid
        self.source_terms, self.implicit_source_term = list(source_terms), implicit_source_term
        if numericalflux.eq is not equation:
            raise ArgumentError("CentralUpwind was built for a different equation")
        # is_compatible(eq, grid) system.jl:81-83
        if isinstance(equation.B, BottomTopography2D):
            want = (grid.totalcells[1] + 1, grid.totalcells[0] + 1)
            if equation.B.B.shape != want:
                raise DomainError(f"Bottom topography should be of size {want[::-1]} but got {equation.B.B.shape[::-1]}")
        # is_compatible(eq, source_terms) equation.jl:74-84
        if not equation.pure and not is_zero(equation.B):
            if not any(isinstance(s, SourceTermBottom) for s in self.source_terms):
                raise ArgumentError("Found non-zero bottom topography in equation, but no corresponding source "
                                    "term. Did you forget to add a source term to your system?")
        if isinstance(reconstruction, LinearReconstruction) and grid.gc < 2:
            raise ArgumentError("LinearReconstruction needs ghostcells > 1")   # @assert, linearreconstruction.jl:59


# ---- simulator.jl --------------------------------------------------------------------------------------
class _InteriorState:
    """current_interior_state(sim): .h/.hu/.hv numpy views of a host copy (interior_volume_variable.jl:168-176).
    Note: `.h` stores w = h + B for bathymetry-carrying equations, exactly like the reference."""

    def __init__(self, w, hu, hv=None):
        self.h, self.hu, self.hv = w, hu, hv

    def __iter__(self):
        return iter((self.h, self.hu) if self.hv is None else (self.h, self.hu, self.hv))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Simulator:
    """Simulator(backend, system, timestepper, grid; cfl=0.25, t0=0.0) (simulator.jl:60-91).
    Owns one device-resident simulation (an opaque swe_b200_handle)."""

    def __init__(self, backend, system, timestepper, grid, cfl=0.25, t0=0.0,
                 rank: int = 0, nranks: int = 1, global_ny: Optional[int] = None, y_offset: int = 0):
        self.dim = getattr(grid, "dim", 2)
        self.grid_1d = None
        if self.dim == 1:
            self.grid_1d = grid
            system, grid = _embed_1d(backend, system, grid)
        self.backend, self.system, self.timestepper, self.grid = backend, system, timestepper, grid
        eq = system.equation
        L = _lib.lib()
        c = _lib.Config()
        c.abi_version = _lib.ABI_VERSION
        self.dtype = np.dtype(backend.realtype)
        c.dtype = 0 if self.dtype == np.float32 else 1
        c.nx, c.ny, c.gc = grid.nx, grid.ny, grid.gc
        c.bc = grid.boundary.code
        c.eq_kind = 1 if eq.pure else 0
        c.recon = 1 if isinstance(system.reconstruction, LinearReconstruction) else 0
        c.stepper = timestepper.code
        fr = system.implicit_source_term
        c.friction = _FRICTION[fr.friction_function] if fr is not None else 0
        c.Cz = fr.Cz if fr is not None else 0.0
        has_B = isinstance(eq.B, BottomTopography2D) and not eq.pure
        zero_B = has_B and is_zero(eq.B)
        # an all-zero BottomTopography2D is "compatible" without SourceTermBottom in the reference
        # (equation.jl:75): treat it as ConstantBottomTopography(0.0), same arithmetic results.
        c.has_bathymetry = 1 if (has_B and not zero_B) else 0
        c.const_B = 0.0 if has_B else float(eq.B.B)
        c.has_bottom_source = 1 if any(isinstance(s, SourceTermBottom) for s in system.source_terms) else 0
        order = [s for s in system.source_terms if isinstance(s, (ConstantRain, TimeDependentRain, FunctionalRain,
                                                                  HortonInfiltration, ConstantInfiltration))]
        rains = [s for s in order if isinstance(s, (ConstantRain, TimeDependentRain, FunctionalRain))]
        infils = [s for s in order if isinstance(s, (HortonInfiltration, ConstantInfiltration))]
        if len(rains) > 1 or len(infils) > 1:
            raise B200Error("at most one rain and one infiltration source term are supported")
        self._horton = None
        self._frain, self._frain_t = None, None
        if rains:
            r = rains[0]
            if isinstance(r, ConstantRain):
                c.rain_kind, c.rain_rate = 1, r.rain_rate
            elif isinstance(r, FunctionalRain):
                c.rain_kind, self._frain = 3, r
            else:
                if len(r.rain_rates) > 16:
                    raise B200Error("TimeDependentRain table too long (max 16)")
                c.rain_kind, c.n_rain = 2, len(r.rain_rates)
                for q, (a, b) in enumerate(zip(r.rain_rates, r.time)):
                    c.rain_rates[q], c.rain_times[q] = a, b
        if infils:
            f = infils[0]
            if isinstance(f, ConstantInfiltration):
                c.infil_kind, c.infil_const = 1, f.infiltration_rate
            else:
                c.infil_kind, c.f0, c.fc, c.k = 2, f.f0, f.fc, f.k
                self._horton = f
        c.infil_first = 1 if (rains and infils and order.index(infils[0]) < order.index(rains[0])) else 0
        c.strict = 1 if backend.strict else 0
        c.fused = 1 if backend.fused else 0
        c.device = backend.device
        c.rank, c.nranks = rank, nranks
        c.global_ny = global_ny if global_ny is not None else grid.ny
        c.y_offset = y_offset
        for q in range(4):
            c.extent[q] = grid.extent[q]
        rec = system.reconstruction
        c.theta = rec.theta if isinstance(rec, LinearReconstruction) else 1.0
        c.rho, c.g = eq.rho, eq.g
        c.depth_cutoff, c.kappa = eq.depth_cutoff, eq.desingularizing_kappa
        c.cfl, c.t0 = float(cfl), float(t0)
        self.cfg = c
        self.cfl = float(cfl)
        h = C.c_void_p()
        rc = L.swe_b200_create(C.byref(c), C.byref(h))
        if rc != 0:
            _raise(rc, None)
        self._h = h
        self.Nx, self.Ny = grid.totalcells
        if c.has_bathymetry:
            Bd = np.ascontiguousarray(eq.B.B, dtype=self.dtype)
            self._check(L.swe_b200_set_bathymetry(self._h, _ptr(Bd), Bd.size))
        if self._horton is not None:
            f = self._horton.factor
            self._check(L.swe_b200_set_infiltration_factor(self._h, _ptr(f), f.size))
        if self._frain is not None:      # cell_center of ALL cells, ghosts included (grid.jl:117-121)
            gc = grid.gc
            dx = (grid.extent[1] - grid.extent[0]) / grid.nx
            dy = (grid.extent[3] - grid.extent[2]) / c.global_ny
            self._frain_x = grid.extent[0] + dx * (np.arange(1, self.Nx + 1) - 0.5 - gc)
            self._frain_y = grid.extent[2] + dy * (np.arange(1, self.Ny + 1) - 0.5 - gc + y_offset)
            self._refresh_rain(float(t0))

    def set_wrap_bathymetry(self, side: int, corner_rows):
        """y-strips: corner values of the gc image-source cell rows at the far end of the domain (swe_b200_set_wrap_bathymetry)"""
        rows = np.ascontiguousarray(corner_rows, dtype=self.dtype)
        self._check(_lib.lib().swe_b200_set_wrap_bathymetry(self._h, int(side), _ptr(rows), rows.size))

    def _check(self, rc):
        if rc != 0:
            _raise(rc, self._h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().swe_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state -------------------------------------------------------------------------------------
    def set_state_interior(self, w, hu, hv=None):
        if self.dim == 1:          # (h, hu) of the 1D problem -> identical rows, hv = 0
            w, hu = (np.tile(np.asarray(a).reshape(1, -1), (ROWS_1D, 1)) for a in (w, hu))
            hv = np.zeros_like(w)
        arrs = [np.ascontiguousarray(a, dtype=self.dtype) for a in (w, hu, hv)]
        for a in arrs:
            if a.shape != (self.grid.ny, self.grid.nx):
                raise ArgumentError(f"state arrays must have shape {(self.grid.ny, self.grid.nx)}, got {a.shape}")
        self._check(_lib.lib().swe_b200_set_state_interior(self._h, *[_ptr(a) for a in arrs]))

    def get_state_interior(self):
        out = [np.empty((self.grid.ny, self.grid.nx), dtype=self.dtype) for _ in range(3)]
        self._check(_lib.lib().swe_b200_get_state_interior(self._h, *[_ptr(a) for a in out]))
        if self.dim == 1:
            return [out[0][0].copy(), out[1][0].copy()]
        return out

    def get_state_full(self, buffer=0):
        out = np.empty((3, self.Ny, self.Nx), dtype=self.dtype)
        self._check(_lib.lib().swe_b200_get_state_full(self._h, buffer, _ptr(out)))
        return out

    def set_state_full(self, data, buffer=0):
        d = np.ascontiguousarray(data, dtype=self.dtype)
        if d.shape != (3, self.Ny, self.Nx):
            raise ArgumentError("full state must have shape (3, Ny, Nx)")
        self._check(_lib.lib().swe_b200_set_state_full(self._h, buffer, _ptr(d)))

    def _time(self):
        t, dt = C.c_double(), C.c_double()
        self._check(_lib.lib().swe_b200_get_time(self._h, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    @property
    def t(self):
        return self._time()[0]

    @property
    def dt(self):
        return self._time()[1]

    def set_time(self, t, dt=0.0):
        self._check(_lib.lib().swe_b200_set_time(self._h, float(t), float(dt)))

    # -- stepping ---------------------------------------------------------------------------------------
    def _refresh_rain(self, t=None):
        """FunctionalRain: (re-)evaluate the closure at the time of the coming step and upload the values"""
        if self._frain is None:
            return
        t = self.t if t is None else t
        if self._frain_t is None or (self._frain.time_dependent and t != self._frain_t):
            v = self._frain.values(t, self._frain_x, self._frain_y)
            self._check(_lib.lib().swe_b200_set_rain_plane(self._h, _ptr(v), v.size))
            self._frain_t = t

    @property
    def host_loop_required(self) -> bool:
        """a time-dependent FunctionalRain needs the host between steps"""
        return self._frain is not None and self._frain.time_dependent

    def step(self, max_dt=math.inf, advance=True):
        if self.host_loop_required:
            self._refresh_rain()
        self._check(_lib.lib().swe_b200_step(self._h, float(max_dt), 1 if advance else 0))

    def run_steps(self, n, maximum_timestep=None):
        if self.host_loop_required:
            for _ in range(int(n)):
                self.step(maximum_timestep if maximum_timestep else math.inf)
            return
        self._check(_lib.lib().swe_b200_run_steps(self._h, int(n), float(maximum_timestep or 0.0)))

    def run_steps_timed(self, n, maximum_timestep=None):
        tm = _lib.Timing()
        self._check(_lib.lib().swe_b200_run_steps_timed(self._h, int(n), float(maximum_timestep or 0.0), C.byref(tm)))
        return dict(total_ms=tm.total_ms, stage_kernel_ms=tm.stage_kernel_ms, launches=tm.launches,
                    stage_launches=tm.stage_launches, stage1_ms=tm.stage1_ms, stage2_ms=tm.stage2_ms)

    def sync(self):
        self._check(_lib.lib().swe_b200_sync(self._h))

    def launch_count(self):
        n = C.c_int64()
        self._check(_lib.lib().swe_b200_launch_count(self._h, C.byref(n)))
        return n.value

    KERNEL_FAMILIES = {0: "per-operator", 1: "fused-generic", 2: "fused-x2", 3: "fused-d1"}

    def kernel_family(self) -> str:
        """which kernel family serves step / run_steps / simulate_to_time on this handle (swe_b200_kernel_family)"""
        rc = _lib.lib().swe_b200_kernel_family(self._h)
        if rc < 0:
            self._check(rc)
        return self.KERNEL_FAMILIES[rc]

    def compare_states(self, other, buffer=0, tol=None):
        """Per-cell comparison with another simulator's state ON THE DEVICE (same grid / dtype / device): dict with
        the rel-L-inf errors of SURVEY 8c (w against max|w|, momenta against max(|hu|,|hv|) of `other`), the number of
        cells whose wet flag differs, the wet-cell counts and -- with `tol` -- the number of cells per field whose
        error exceeds tol x the field's scale."""
        L = _lib.lib()
        out = (C.c_double * 13)()
        self._check(L.swe_b200_compare_states(self._h, other._h, int(buffer), None, out))
        o = list(out)
        dw = max(o[3], 1e-300)
        dm = max(o[4], o[5], 1e-300)
        res = dict(rel_linf=[o[0] / dw, o[1] / dm, o[2] / dm], max_abs_diff=o[0:3], max_abs_ref=o[3:6],
                   mask_mismatch=int(o[6]), wet=int(o[7]), wet_other=int(o[8]), nan_cells=int(o[9]))
        if tol is not None:
            thr = (C.c_double * 3)(tol * dw, tol * dm, tol * dm)
            self._check(L.swe_b200_compare_states(self._h, other._h, int(buffer), thr, out))
            res["cells_over_tol"] = [int(out[10]), int(out[11]), int(out[12])]
        return res

    # -- fine-grained operators ---------------------------------------------------------------------------
    def update_bc(self, buffer=0):
        self._check(_lib.lib().swe_b200_update_bc(self._h, buffer))

    def reconstruct(self, direction, buffer=0):
        self._check(_lib.lib().swe_b200_reconstruct(self._h, direction, buffer))

    def get_faces(self):
        l = np.empty((3, self.Ny, self.Nx), dtype=self.dtype)
        r = np.empty((3, self.Ny, self.Nx), dtype=self.dtype)
        self._check(_lib.lib().swe_b200_get_faces(self._h, _ptr(l), _ptr(r)))
        return l, r

    def add_time_derivative(self, in_buffer=0, out_buffer=1, t=0.0):
        ms = (C.c_double * 2)()
        self._check(_lib.lib().swe_b200_add_time_derivative(self._h, in_buffer, out_buffer, float(t), ms))
        return ms[0], ms[1]

    def get_wavespeeds(self):
        w = np.empty((self.Ny, self.Nx), dtype=self.dtype)
        self._check(_lib.lib().swe_b200_get_wavespeeds(self._h, _ptr(w)))
        return w

    # -- asynchronous snapshots (writer callbacks) ---------------------------------------------------
    def snapshot_async(self, out=None):
        """Start copying the current interior state to page-locked host arrays while the simulation goes on
        (swe_b200_snapshot_begin); returns a Snapshot whose .wait() gives (w, hu, hv) -- (h, hu) for a 1D problem.
        `out` may be the arrays of an earlier Snapshot (reused, no new allocation)."""
        if out is None:
            out = [pinned_empty((self.grid.ny, self.grid.nx), self.dtype) for _ in range(3)]
        self._check(_lib.lib().swe_b200_snapshot_begin(self._h, *[_ptr(a) for a in out]))
        return Snapshot(self, out, self.t)

    DIAG_FIELDS = ("sum_h", "sum_hu", "sum_hv", "max_h", "min_h", "wet_cells", "max_momentum", "sum_w")

    def diagnostics(self, buffer=0) -> dict:
        """Scalar reductions over the interior cells, computed on the device (64 bytes come back instead of
        collect(state)): what TotalWaterVolume-style callbacks need (examples/example_tools.jl:86-90)."""
        out = (C.c_double * 8)()
        self._check(_lib.lib().swe_b200_diagnostics(self._h, buffer, out))
        d = dict(zip(self.DIAG_FIELDS, [float(v) for v in out]))
        d["wet_cells"] = int(d["wet_cells"])
        if self.dim == 1:          # a strip of ROWS_1D identical rows stands for the 1D domain
            for k in ("sum_h", "sum_hu", "sum_hv", "sum_w"):
                d[k] /= ROWS_1D
            d["wet_cells"] //= ROWS_1D
        return d

    def implicit_substep(self, out_buffer, prev_buffer, dt):
        self._check(_lib.lib().swe_b200_implicit_substep(self._h, out_buffer, prev_buffer, float(dt)))

    def post_proc_substep(self, buffer):
        self._check(_lib.lib().swe_b200_post_proc_substep(self._h, buffer))

    def enumerate(self, kind, direction=0, ghostcells=None):
        L = _lib.lib()
        n = C.c_int64()
        g = self.grid.gc if ghostcells is None else int(ghostcells)
        if kind == "cells":
            call = lambda out: L.swe_b200_enumerate_cells(self._h, out, C.byref(n))
            per = 2
        elif kind == "inner":
            call = lambda out: L.swe_b200_enumerate_inner_cells(self._h, direction, g, out, C.byref(n))
            per = 6
        else:
            call = lambda out: L.swe_b200_enumerate_ghost_cells(self._h, direction, out, C.byref(n))
            per = 2
        self._check(call(None))
        out = np.empty((n.value, per), dtype=np.int64)
        if n.value:
            self._check(call(_ptr(out)))
        return out


def set_current_state(simulator: Simulator, new_state):
    """set_current_state!(sim, state) (simulator.jl:119-139): interior values, then update_bc!."""
    simulator.set_state_interior(*new_state)


def current_interior_state(simulator: Simulator) -> _InteriorState:
    return _InteriorState(*simulator.get_state_interior())


def current_state(simulator: Simulator):
    return simulator.get_state_full(0)


def current_time(simulator: Simulator):
    return simulator.t


def current_timestep(simulator: Simulator):
    return simulator.dt


def perform_step(simulator: Simulator, max_dt):
    """perform_step!(sim, max_dt) (simulator.jl:145-177); does not advance t."""
    simulator.step(max_dt, advance=False)


def update_bc(simulator: Simulator, buffer=0):
    simulator.update_bc(buffer)


def simulate_to_time(simulator: Simulator, endtime, match_endtime=True, callback=None, show_progress=False,
                     maximum_timestep=None):
    """simulate_to_time (simulator.jl:179-212).  Without a callback the loop runs device-resident
    (dt and t never visit the host between steps); with one, the reference's per-step host loop is kept."""
    L = _lib.lib()
    if callback is None and simulator.host_loop_required:
        callback = lambda t, s: None
    if callback is None:
        n = C.c_int64()
        simulator._check(L.swe_b200_simulate_to_time(simulator._h, float(endtime), 1 if match_endtime else 0,
                                                     float(maximum_timestep or 0.0), C.byref(n)))
        return n.value
    if isinstance(callback, (list, tuple)):
        cbs = list(callback)
        callback = lambda t, s: [cb(t, s) for cb in cbs]     # MultipleCallbacks, callbacks.jl:74-82
    n = 0
    t = simulator.t
    while t < endtime:
        max_dt = endtime - t if match_endtime else math.inf
        if maximum_timestep is not None:
            max_dt = min(max_dt, maximum_timestep)
        simulator.step(max_dt, advance=True)
        t = simulator.t
        n += 1
        callback(t, simulator)
    return n


class IntervalWriter:
    """callbacks.jl:40-55: call `writer(t, simulator)` every `step` units of simulated time."""

    def __init__(self, step: float, writer: Callable):
        self.step, self.writer, self.current_t = float(step), writer, 0.0

    def __call__(self, t, simulator):
        dt = simulator.dt
        if t + dt >= self.current_t:
            self.writer(t, simulator)
            self.current_t += self.step


class _PinnedBlock:
    """a cudaHostAlloc block exposed through the array interface: numpy views keep it alive, __del__ frees it"""

    def __init__(self, nbytes):
        self.ptr = _lib.lib().swe_b200_host_alloc(int(nbytes))
        if not self.ptr:
            raise B200Error("cudaHostAlloc failed")
        self.__array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(self.ptr), False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().swe_b200_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """page-locked host array (swe_b200_host_alloc): device<->host copies of it run asynchronously at PCIe speed"""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    block = _PinnedBlock(max(count * dtype.itemsize, 1))
    return np.asarray(block)[:count * dtype.itemsize].view(dtype).reshape(shape)


class Snapshot:
    """state copy in flight (Simulator.snapshot_async): .wait() blocks until the host arrays are complete"""

    def __init__(self, simulator, arrays, t):
        self.simulator, self.arrays, self.t, self._done = simulator, arrays, t, False

    def wait(self):
        if not self._done:
            self.simulator._check(_lib.lib().swe_b200_snapshot_wait(self.simulator._h))
            self._done = True
        if self.simulator.dim == 1:
            return [self.arrays[0][0], self.arrays[1][0]]
        return self.arrays


class AsyncIntervalWriter:
    """IntervalWriter (callbacks.jl:40-55) whose writer receives a finished host copy of the state instead of the
    simulator: at every interval a snapshot is started and the PREVIOUS one is handed to `writer(t, (w, hu, hv))`, so
    the device->host copy overlaps the following steps.  Call .finish() after simulate_to_time for the last one."""

    def __init__(self, step: float, writer: Callable):
        self.step, self.writer, self.current_t, self._pending, self._spare = float(step), writer, 0.0, None, None

    def _flush(self):
        if self._pending is not None:
            snap = self._pending
            self.writer(snap.t, snap.wait())
            self._spare, self._pending = snap.arrays, None

    def __call__(self, t, simulator):
        if t + simulator.dt >= self.current_t:
            self._flush()
            self._pending = simulator.snapshot_async(self._spare)
            self._spare = None
            self.current_t += self.step

    def finish(self):
        self._flush()


class TotalWaterVolume:
    """examples/example_tools.jl:62-90: callback recording sum(w - B_cell) over the interior at every call.
    The reference collects the whole state to the host; here the sum is a device reduction
    (swe_b200_diagnostics).  `reduce` combines the per-strip sums of a multi-GPU run (e.g. an all-reduce)."""

    def __init__(self, reduce: Optional[Callable] = None):
        self.times, self.water_volume, self.reduce = [], [], reduce

    def __call__(self, t, simulator):
        v = simulator.diagnostics()["sum_h"]
        self.water_volume.append(self.reduce(v) if self.reduce else v)
        self.times.append(t)


# ---- the backend kernel surface ---------------------------------------------------------------------
# for_each_cell(f, backend, grid, y...), for_each_inner_cell(f, backend, grid, direction, y...; ghostcells),
# for_each_ghost_cell(f, backend, grid, direction, y...) (kernel_abstractions.jl:73, 83, 129).  A precompiled
# backend cannot execute an arbitrary closure on the device: the index space is enumerated by a device kernel
# (swe_b200_enumerate_grid -- no simulator needed, like the reference's signature) and `f` is applied on the host
# to each 1-based index with the trailing arguments y...  The hot-path CALLERS of this surface (reconstruct!,
# compute_flux!, do_substep!, update_bc!, ...) never come here: they have fused CUDA implementations.
def _enumerate_grid(kind, backend, grid, direction=0, ghostcells=0):
    L = _lib.lib()
    n = C.c_int64()
    args = (kind, grid.nx, grid.ny, grid.gc, int(direction), int(ghostcells), backend.device)
    rc = L.swe_b200_enumerate_grid(*args, None, C.byref(n))
    if rc:
        _raise(rc)
    out = np.empty((n.value, 6 if kind == 1 else 2), dtype=np.int64)
    if n.value:
        rc = L.swe_b200_enumerate_grid(*args, _ptr(out), C.byref(n))
        if rc:
            _raise(rc)
    return out


def for_each_cell(f, backend, grid, *y):
    """for_each_cell(f, backend, grid, y...) (kernel_abstractions.jl:123-131): f(index, y...), 1-based (i, j)."""
    for i, j in _enumerate_grid(0, backend, grid):
        f((int(i), int(j)), *y)


def for_each_inner_cell(f, backend, grid, direction, *y, ghostcells=None):
    """for_each_inner_cell(f, backend, grid, direction, y...; ghostcells=grid.ghostcells[direction])
    (kernel_abstractions.jl:66-75): f(left, middle, right, y...) with 1-based index pairs."""
    g = grid.gc if ghostcells is None else int(ghostcells)
    for r in _enumerate_grid(1, backend, grid, direction, g):
        f((int(r[0]), int(r[1])), (int(r[2]), int(r[3])), (int(r[4]), int(r[5])), *y)


def for_each_ghost_cell(f, backend, grid, direction, *y):
    """for_each_ghost_cell(f, backend, grid, direction, y...) (kernel_abstractions.jl:77-85): f(ghostcell, y...)."""
    for i, j in _enumerate_grid(2, backend, grid, direction):
        f((int(i), int(j)), *y)


# ---- convenience for tests / bench: build a Simulator from a case dict (tests/cases.py) -----------------
_BC = {"periodic": PeriodicBC, "wall": WallBC, "neumann": NeumannBC}


def simulator_from_case(case: dict, strict: bool = False, fused: bool = True, device: int = 0,
                        rank: int = 0, nranks: int = 1, global_ny=None, y_offset: int = 0) -> Simulator:
    dim = case.get("dim", 2)
    dtype = np.dtype(case.get("dtype", "float64"))
    backend = make_b200_backend(dtype.type, strict=strict, fused=fused, device=device)
    if dim == 1:
        ext = case.get("extent", [0.0, 1.0])
        grid = CartesianGrid(case["nx"], gc=case.get("gc", 2), boundary=_BC[case.get("bc", "periodic")](),
                             extent=[ext[0], ext[1]])
    else:
        ext = case.get("extent", [0.0, 1.0, 0.0, 1.0])
        grid = CartesianGrid(case["nx"], case["ny"], gc=case.get("gc", 2), boundary=_BC[case.get("bc", "periodic")](),
                             extent=[[ext[0], ext[1]], [ext[2], ext[3]]])
    B = case.get("B")
    if case.get("pure", False):
        eq = (ShallowWaterEquations1DPure if dim == 1 else ShallowWaterEquationsPure)(case.get("rho", 1.0), case.get("g", 9.81))
    else:
        if B is None:
            bt = ConstantBottomTopography()
        elif np.isscalar(B):
            bt = ConstantBottomTopography(B)
        else:
            bt = (BottomTopography1D if dim == 1 else BottomTopography2D)(B, backend, grid)
        eq = (ShallowWaterEquations1D if dim == 1 else ShallowWaterEquations)(
            bt, rho=case.get("rho", 1.0), g=case.get("g", 9.81), depth_cutoff=case.get("depth_cutoff", 1e-5),
            desingularizing_kappa=case.get("kappa", 1e-5))
    rec = LinearReconstruction(case.get("theta", 1.2)) if case.get("recon", "linear") == "linear" else NoReconstruction()
    sources = []
    rain, infil = case.get("rain"), case.get("infiltration")
    infil_term = rain_term = None
    if infil is not None:
        infil_term = ConstantInfiltration(infil) if np.isscalar(infil) else HortonInfiltration(
            grid, backend, factor=infil.get("factor", 1.0), fc=infil.get("fc", 3.272e-5),
            f0=infil.get("f0", 1.977e-4), k=infil.get("k", 2.43e-3))
    if rain is not None:
        if isinstance(rain, dict):
            rain_term = FunctionalRain(rain["function"], grid, time_dependent=rain.get("time_dependent", True))
        else:
            rain_term = ConstantRain(rain) if np.isscalar(rain) else TimeDependentRain(rain[0], rain[1])
    pair = [infil_term, rain_term] if case.get("infil_first", True) else [rain_term, infil_term]
    sources += [s for s in pair if s is not None]
    if case.get("bottom_source", False):
        sources.append(SourceTermBottom())
    fr = case.get("friction")
    friction = ImplicitFriction(Cz=case.get("Cz", 0.03 ** 2), friction_function=fr) if fr else None
    system = ConservedSystem(backend, rec, CentralUpwind(eq), eq, grid, sources, friction)
    stepper = RungeKutta2() if case.get("stepper", "rk2") == "rk2" else ForwardEulerStepper()
    sim = Simulator(backend, system, stepper, grid, cfl=case.get("cfl", 0.25), t0=case.get("t0", 0.0),
                    rank=rank, nranks=nranks, global_ny=global_ny, y_offset=y_offset)
    for side, key in ((0, "wrap_B_lo"), (1, "wrap_B_hi")):        # y-strips with wrap-around ghost rows (multigpu.strip_case)
        if case.get(key) is not None:
            sim.set_wrap_bathymetry(side, case[key])
    return sim
