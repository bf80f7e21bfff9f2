"""Builds one host-emulation translation unit of tests/cuda_emu/ (see tests/cuda_emu/emu.h) into a shared library.
TEST INFRASTRUCTURE ONLY: the product binds libgemmgan_sm100a.so and nothing else."""
import ctypes as C
import hashlib
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
CUDA_LIB = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/lib64"


def build(unit: str, outdir, cudart: bool = False) -> C.CDLL:
    """unit = 'evalmetrics' -> tests/cuda_emu/emu_evalmetrics.cpp. GEMMGAN_EMU_ASAN=1 (with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0) adds AddressSanitizer, so the kernels'
    global-memory accesses are checked against the redzones of the numpy / torch allocations."""
    gxx = shutil.which("g++")
    if gxx is None or not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    src = os.path.join(ROOT, "tests", "cuda_emu", f"emu_{unit}.cpp")
    out = os.path.join(str(outdir), f"lib{unit}_emu.so")
    extra = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if os.environ.get("GEMMGAN_EMU_ASAN") == "1" else []
    # builds are cached by the content of every source that can reach the translation unit (several test modules use
    # the engine build, which takes ~25 s to compile)
    h = hashlib.sha256(" ".join(extra + [unit, str(cudart)]).encode())
    for d in (os.path.join(ROOT, "tests", "cuda_emu"), os.path.join(ROOT, "gemmgan_b200", "csrc"), os.path.join(ROOT, "include")):
        for fn in sorted(os.listdir(d)):
            if fn.endswith((".h", ".cuh", ".cu", ".cpp")):
                h.update(fn.encode())
                h.update(open(os.path.join(d, fn), "rb").read())
    cache_dir = os.path.join(ROOT, "tests", "cuda_emu", "build")
    os.makedirs(cache_dir, exist_ok=True)
    cached = os.path.join(cache_dir, f"lib{unit}_emu_{h.hexdigest()[:16]}.so")
    if os.path.exists(cached):
        shutil.copyfile(cached, out)
        L = C.CDLL(out)
        L.gg_last_error.restype = C.c_char_p
        return L
    subprocess.check_call([gxx, "-std=c++20", "-O1", "-fno-extern-tls-init", *extra, "-shared", "-fPIC", "-pthread", "-I", CUDA_INC,
                           "-I", os.path.join(ROOT, "include"), src, "-o", out] + (
        # host-side runtime symbols some files reference but the emulated paths never call
        ["-L", CUDA_LIB, "-Wl,-rpath," + CUDA_LIB, "-lcudart"] if cudart else []))
    tmp = cached + f".{os.getpid()}.tmp"
    shutil.copyfile(out, tmp)
    os.replace(tmp, cached)
    L = C.CDLL(out)
    L.gg_last_error.restype = C.c_char_p
    return L
