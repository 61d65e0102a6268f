"""qwen3.c_b200 -- B200-native forward hot path for qwen3.c checkpoints.

The product is the C-ABI library `lib/libqwen3.so` built from `csrc/` (host C + sm_100a
CUDA). This Python package only holds what surrounds it: the build driver, a ctypes
mirror of the reference's operator interface (include/forward.h, q8.h, model.h) and the
synthetic-checkpoint writer used by tests and bench.py.
"""
from . import binding, build, checkpoint, tp  # noqa: F401
from .binding import B200Model, QwenLib  # noqa: F401
