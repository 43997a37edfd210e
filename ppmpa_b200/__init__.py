"""ppmpa_b200 -- B200-native progressive photon mapping hot path (eijian/ppmpa).

The product is libppm_b200.so (CUDA sm_100a kernels + the C ABI of
include/ppm.h); this package is the thin host mirror used by tests and bench.
"""
from . import _capi  # noqa: F401  (raises ImportError if the .so is missing)
from ._capi import (FILTER_CONE, FILTER_GAUSS, FILTER_NONE, PHOTON_DTYPE, WL_BLUE, WL_GREEN, WL_RED)  # noqa: F401
from .engine import (Engine, PPMError, Scene, format_f64, radius_schedule, read_camera, read_scene)  # noqa: F401
from .parallel import passes_for_rank  # noqa: F401
