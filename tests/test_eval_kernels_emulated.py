"""The evaluation-metric kernels (gemmgan_b200/csrc/evalmetrics.cu) checked WITHOUT a GPU.

tests/cuda_emu/emu_evalmetrics.cpp compiles the unchanged .cu with g++ and runs every CUDA thread as a fiber
(counting barriers for __syncthreads, slot exchange for warp shuffles; CTAs spread over the host cores). That build exports the same C-ABI entry points, so the
host mirror gemmgan_b200/evalmetrics.py is run against it end to end by pointing three module-level hooks (device,
stream, library) at the emulation. This checks index arithmetic, tile edges, reductions, rank selection and the
argument validation; it says nothing about performance and is no substitute for tests/test_gpu_zeval.py on a B200.

The emulation is test infrastructure: the product never loads it (it only binds libgemmgan_sm100a.so).
"""
import pytest
import torch

import emu_build
from gemmgan_b200 import _abi_decl, _lib
from gemmgan_b200 import evalmetrics as em
import eval_cases
from eval_cases import *  # noqa: F401,F403  (the shared test functions)



@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("evalmetrics", tmp_path_factory.mktemp("cuda_emu"))
    _abi_decl.declare_evalmetrics(L)
    return L


@pytest.fixture()
def host(emu, monkeypatch):
    """gemmgan_b200.evalmetrics with its device / stream / library hooks pointed at the emulation."""
    monkeypatch.setattr(em, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(em, "_stream", lambda: None)
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setitem(eval_cases.DEV, "device", "cpu")
    return em
