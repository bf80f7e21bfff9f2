"""Test infrastructure: runs the drop-in trainer classes (gemmgan_b200/trainer.py::TrainerBase) on the CPU against the
host-emulated engine (tests/cuda_emu). `apply(setattr, lib)` redirects the few places where the trainer touches CUDA:
the library handle, the device, events, pinned memory, the copy stream; steps run eagerly (no CUDA graphs), on one lane,
with the CUDA-core fp32 GEMM path (the host build has no tcgen05 kernels). Pass `monkeypatch.setattr` inside pytest,
or the built-in `setattr` in a spawned worker process. The product never imports this file."""
import contextlib

import torch

from gemmgan_b200 import _lib, datasets, runtime, trainer


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


def apply(setattr_, lib, dropout_p=0.0):
    setattr_(_lib, "lib", lambda: lib)
    setattr_(_lib, "require_device", lambda dev=0: None)
    setattr_(_lib, "require_cuda_tensor_device", lambda dev, what: None)
    setattr_(runtime, "_stream", lambda: None)
    setattr_(datasets, "_default_device", lambda: torch.device("cpu"))     # DeviceResidentLoader: "HBM" = host memory
    setattr_(datasets, "_current_stream", lambda: None)
    setattr_(torch.cuda, "device", lambda d: contextlib.nullcontext())
    setattr_(torch.cuda, "Event", _Event)
    setattr_(torch.Tensor, "pin_memory", lambda self: self)
    orig_engine = runtime.Engine.__init__

    def simt(self, *a, **kw):
        kw["gemm_impl"] = _lib.IMPL_SIMT_F32
        orig_engine(self, *a, **kw)
        self.set_lanes(False)
    setattr_(runtime.Engine, "__init__", simt)
    orig_common = trainer.TrainerBase._init_common

    def common(self, *a, **kw):
        # only around the constructor's device check: torch.optim (the oracle's optimizers) asks torch.cuda too
        saved = torch.cuda.is_available, torch.cuda.current_device
        torch.cuda.is_available, torch.cuda.current_device = (lambda: True), (lambda: 0)
        try:
            orig_common(self, *a, **kw)
        finally:
            torch.cuda.is_available, torch.cuda.current_device = saved
        self.device = torch.device("cpu")
        self.use_cuda_graphs = False
        if dropout_p is not None:      # (the masks are the engine's own stream; dropout-on parity is in the GPU suite)
            self.dropout_p = dropout_p
    setattr_(trainer.TrainerBase, "_init_common", common)
    setattr_(trainer.TrainerBase, "prefetch", lambda self, *t: None)   # a copy stream: GPU only
    return trainer
