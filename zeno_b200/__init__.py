"""zeno_b200 -- B200-native (sm_100a) implementation of the hot path of Zeno's FastFLIP solver:
particle binning, PIC/FLIP particle<->grid transfers on sparse 8^3 leaves and the matrix-free
MGPCG pressure projection, behind a C ABI (include/flipb200.h, zeno_b200/libflipb200.so) that the
Zeno node shims in zeno_b200/plugin/ call. This package holds the CUDA sources (csrc/), the ctypes
driver used by tests and bench (abi.py) and the synthetic scene generator (scenes.py)."""
from . import abi, scenes  # noqa: F401

__all__ = ["abi", "scenes"]
