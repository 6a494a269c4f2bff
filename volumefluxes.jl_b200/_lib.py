"""ctypes binding of csrc/libswe_b200.so (include/swe_b200.h).  There is NO fallback: if the CUDA
library is missing or fails to load, importing the compute API raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(os.environ.get("SWE_B200_LIB", _CSRC / "libswe_b200.so"))   # override: A/B builds

ABI_VERSION = 1
OK, EINVAL, EDOMAIN, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4


class Config(C.Structure):
    """swe_b200_config (include/swe_b200.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("dtype", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("gc", C.c_int32),
        ("bc", C.c_int32), ("eq_kind", C.c_int32), ("recon", C.c_int32), ("stepper", C.c_int32),
        ("friction", C.c_int32), ("has_bathymetry", C.c_int32), ("has_bottom_source", C.c_int32),
        ("rain_kind", C.c_int32), ("infil_kind", C.c_int32), ("infil_first", C.c_int32),
        ("n_rain", C.c_int32), ("strict", C.c_int32), ("fused", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("global_ny", C.c_int32), ("y_offset", C.c_int32),
        ("reserved", C.c_int32 * 3),
        ("extent", C.c_double * 4),
        ("theta", C.c_double), ("rho", C.c_double), ("g", C.c_double), ("depth_cutoff", C.c_double),
        ("kappa", C.c_double), ("cfl", C.c_double), ("Cz", C.c_double), ("const_B", C.c_double),
        ("t0", C.c_double),
        ("rain_rate", C.c_double), ("infil_const", C.c_double), ("f0", C.c_double), ("fc", C.c_double),
        ("k", C.c_double),
        ("rain_rates", C.c_double * 16), ("rain_times", C.c_double * 16),
    ]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("stage_kernel_ms", C.c_double),
                ("launches", C.c_int64), ("stage_launches", C.c_int64),
                ("stage1_ms", C.c_double), ("stage2_ms", C.c_double)]


# every symbol include/swe_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pi64, _pd = C.POINTER(C.c_int64), C.POINTER(C.c_double)
SYMBOLS = {
    "swe_b200_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "swe_b200_destroy": (_i, [_vp]),
    "swe_b200_last_error": (C.c_char_p, [_vp]),
    "swe_b200_sync": (_i, [_vp]),
    "swe_b200_set_bathymetry": (_i, [_vp, _vp, _i64]),
    "swe_b200_set_rain_plane": (_i, [_vp, _vp, _i64]),
    "swe_b200_set_wrap_bathymetry": (_i, [_vp, _i, _vp, _i64]),
    "swe_b200_set_infiltration_factor": (_i, [_vp, _vp, _i64]),
    "swe_b200_set_state_interior": (_i, [_vp, _vp, _vp, _vp]),
    "swe_b200_get_state_interior": (_i, [_vp, _vp, _vp, _vp]),
    "swe_b200_set_state_full": (_i, [_vp, _i, _vp]),
    "swe_b200_get_state_full": (_i, [_vp, _i, _vp]),
    "swe_b200_set_time": (_i, [_vp, _d, _d]),
    "swe_b200_get_time": (_i, [_vp, _pd, _pd]),
    "swe_b200_step": (_i, [_vp, _d, _i]),
    "swe_b200_run_steps": (_i, [_vp, _i64, _d]),
    "swe_b200_simulate_to_time": (_i, [_vp, _d, _i, _d, _pi64]),
    "swe_b200_update_bc": (_i, [_vp, _i]),
    "swe_b200_reconstruct": (_i, [_vp, _i, _i]),
    "swe_b200_get_faces": (_i, [_vp, _vp, _vp]),
    "swe_b200_add_time_derivative": (_i, [_vp, _i, _i, _d, _pd]),
    "swe_b200_get_wavespeeds": (_i, [_vp, _vp]),
    "swe_b200_diagnostics": (_i, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "swe_b200_snapshot_begin": (_i, [_vp, _vp, _vp, _vp]),
    "swe_b200_snapshot_wait": (_i, [_vp]),
    "swe_b200_host_alloc": (_vp, [_i64]),
    "swe_b200_host_free": (_i, [_vp]),
    "swe_b200_implicit_substep": (_i, [_vp, _i, _i, _d]),
    "swe_b200_post_proc_substep": (_i, [_vp, _i]),
    "swe_b200_enumerate_cells": (_i, [_vp, _vp, _pi64]),
    "swe_b200_enumerate_inner_cells": (_i, [_vp, _i, _i, _vp, _pi64]),
    "swe_b200_enumerate_ghost_cells": (_i, [_vp, _i, _vp, _pi64]),
    "swe_b200_enumerate_grid": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _pi64]),
    "swe_b200_get_ipc_handle": (_i, [_vp, _vp]),
    "swe_b200_open_peers": (_i, [_vp, _vp]),
    "swe_b200_close_peers": (_i, [_vp]),
    "swe_b200_run_steps_timed": (_i, [_vp, _i64, _d, C.POINTER(Timing)]),
    "swe_b200_launch_count": (_i, [_vp, _pi64]),
    "swe_b200_kernel_family": (_i, [_vp]),
    "swe_b200_compare_states": (_i, [_vp, _vp, _i, _pd, _pd]),
}


def build(verbose: bool = False) -> Path:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", str(_CSRC), "-j4"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libswe_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library (no compute happens at load time; works without a GPU)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for this path)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)     # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
