"""ninterp_b200 — B200-native batch engine for ninterp's `Interpolator::interpolate`.

Host-side mirror of the reference crate's API (see interpolator.py) over a C-ABI CUDA library
(include/ninterp_cuda.h, csrc/).  Importing this package never computes on the CPU: if
libninterp_cuda.so is missing, using any interpolator raises.
"""
from . import strategy
from .error import EngineError, InterpolateError, ReferencePanic, ValidateError
from .extrapolate import Extrapolate
from .interpolator import Interp0D, Interp1D, Interp2D, Interp3D, InterpND, InterpolatorEnum
from .multi import ReplicatedInterpolator
from .serde import from_json, to_json

__all__ = [
    "strategy", "Extrapolate", "Interp0D", "Interp1D", "Interp2D", "Interp3D", "InterpND", "InterpolatorEnum",
    "ValidateError", "InterpolateError", "ReferencePanic", "EngineError", "ReplicatedInterpolator", "from_json", "to_json",
]
__version__ = "0.1.0"
