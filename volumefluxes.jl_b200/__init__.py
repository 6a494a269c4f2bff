"""volumefluxes.jl_b200 -- B200-native backend for the 2D shallow-water finite-volume step of
sintefmath/VolumeFluxes.jl (module SinFVM).

Host-side mirror of the reference's operator interface (Simulator / ShallowWaterEquations /
CentralUpwind / LinearReconstruction / RungeKutta2 ...) over the C-ABI library
`csrc/libswe_b200.so` (include/swe_b200.h).  There is no CPU fallback: every compute call
goes through the CUDA library and fails loudly if it is missing.
"""
from . import synthetic  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
from . import multigpu  # noqa: F401
from . import workloads  # noqa: F401
