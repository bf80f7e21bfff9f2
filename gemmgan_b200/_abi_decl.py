"""argtypes for the entry points of include/gemmgan.h beyond gg_gemm_bf16."""
from __future__ import annotations

import ctypes as C


def declare(L: C.CDLL) -> None:
    pass
