"""ctypes binding of libgemmgan_sm100a.so (include/gemmgan.h).

There is deliberately no fallback: if the library is missing or the device is not sm_100 the
import / call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgemmgan_sm100a.so"

GG_OK = 0
ABI_VERSION = 4   # must equal GG_ABI_VERSION of include/gemmgan.h: bumped with every struct / signature change
ACT_NONE, ACT_LEAKY, ACT_FILM = 0, 1, 2
IMPL_TCGEN05, IMPL_SIMT_F32 = 0, 1


class GGError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("pre", C.c_void_p),
        ("pre_ld", C.c_int64),
        ("pre_f32", C.c_int32),
        ("act", C.c_int32),
        ("slope", C.c_float),
        ("drop_p", C.c_float),
        ("rng", C.c_void_p),
        ("site", C.c_uint32),
        ("mask", C.c_void_p),
        ("mask_ld", C.c_int64),
        ("mask_f32", C.c_int32),
        ("mask_pos", C.c_float),
        ("mask_neg", C.c_float),
        ("res", C.c_void_p),
        ("res_ld", C.c_int64),
        ("res_f32", C.c_int32),
        ("out_bf16", C.c_void_p),
        ("ld_bf16", C.c_int64),
        ("out_f32", C.c_void_p),
        ("ld_f32", C.c_int64),
        ("accum_f32", C.c_int32),
        ("row_div", C.c_int32),
        ("row_mul", C.c_int32),
        ("row_add", C.c_int32),
    ]


class GemmSeg(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("b", C.c_void_p),
        ("lda", C.c_int64),
        ("ldb", C.c_int64),
        ("K", C.c_int32),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("nseg", C.c_int32),
        ("seg", GemmSeg * 2),
        ("a_mn_major", C.c_int32),
        ("b_mn_major", C.c_int32),
        ("epi", Epilogue),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("impl", C.c_int32),
        ("force_splits", C.c_int32),
        ("block_n", C.c_int32),
        ("light", C.c_int32),
        ("pair", C.c_int32),
        ("tf32_operands", C.c_int32),
    ]


_lib = None


def build_if_needed() -> Path:
    if not LIB_PATH.exists() or os.environ.get("GEMMGAN_REBUILD") == "1":
        from . import build as _build

        _build.build()
    return LIB_PATH


def lib() -> C.CDLL:
    """Loads the shared library (building it with nvcc first if it is not there)."""
    global _lib
    if _lib is None:
        path = build_if_needed()
        if not path.exists():
            raise GGError(f"{path} is missing and could not be built; there is no CPU fallback")
        L = C.CDLL(str(path))
        L.gg_last_error.restype = C.c_char_p
        L.gg_abi_version.restype = C.c_int
        got = L.gg_abi_version()
        if got != ABI_VERSION:
            # a stale git-ignored .so after a header change would be called with mismatched ctypes layouts
            raise GGError(f"{path} was built for ABI version {got}, the Python bindings declare {ABI_VERSION}: "
                          "rebuild it (python gemmgan_b200/build.py, or GEMMGAN_REBUILD=1)")
        L.gg_check_device.argtypes = [C.c_int]
        L.gg_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        _declare_rest(L)
        _lib = L
    return _lib


def _declare_rest(L: C.CDLL) -> None:
    """Declares argtypes of the remaining entry points when the library exports them."""
    from . import _abi_decl

    _abi_decl.declare(L)


def check(rc: int) -> None:
    if rc != GG_OK:
        raise GGError(f"libgemmgan_sm100a error {rc}: {lib().gg_last_error().decode()}")


def require_device(dev: int = 0) -> None:
    check(lib().gg_check_device(dev))


def require_cuda_tensor_device(dev, what: str) -> None:
    """Free-standing modules: the parameters must already live on the sm_100a device (no CPU fallback path)."""
    if dev.type != "cuda":
        raise RuntimeError(f"{what} runs on the sm_100a engine: move the module to a CUDA device first "
                           "(there is no PyTorch / CPU fallback path)")
    require_device(dev.index or 0)
