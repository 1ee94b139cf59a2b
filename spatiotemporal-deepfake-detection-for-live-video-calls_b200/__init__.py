"""afb200 — B200-native AltFreezing clip-classification hot path.

Host side (Python) of the C-ABI library in csrc/ (include/afb200.h).  Importing
the package does not load the CUDA library; `afb200.lib()` does, and raises if
the built `libafb200.so` is missing — there is no CPU fallback.
"""
from . import arch, synthetic  # noqa: F401

__all__ = ["arch", "synthetic"]
