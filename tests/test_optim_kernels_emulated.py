"""The optimizer kernels of the hot path (gemmgan_b200/csrc/optim.cu: global gradient norm + clip coefficient, fused
RMSprop / Adam / AdamW update on flat buffers, fp32 -> bf16 weight-shadow refresh) checked WITHOUT a
GPU: the unchanged .cu is compiled for the host (tests/cuda_emu/emu.h) and compared with torch.optim +
clip_grad_norm_ — what the reference calls (src/conditional_gan_cross_attention_with_film.py:320-331, :414-415).
The same comparison runs on the B200 in tests/test_gpu_parity.py::test_optimizer_kernel_matches_torch."""
import ctypes as C

import numpy as np
import pytest
import torch

import emu_build


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("optim", tmp_path_factory.mktemp("cuda_emu"))
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.gg_optim_step.argtypes = [i32, vp, vp, vp, vp, i64, f32, f32, vp, vp, vp, vp]
    L.emu_refresh_shadows.argtypes = [vp, vp, vp, i32, vp]
    return L


def check(L, rc):
    assert rc == 0, L.gg_last_error()


@pytest.mark.parametrize("kind,name", [(0, "rms_prop"), (1, "adam"), (2, "adamw")])
@pytest.mark.parametrize("clip", [0.0, 0.5])
@pytest.mark.parametrize("n", [3, 1024, 10007])          # tails of 3 / 0 / 3 elements behind the float4 body
def test_optimizer_kernel_matches_torch(emu, kind, name, clip, n):
    g0 = torch.Generator().manual_seed(n + kind)
    p = torch.randn(n, generator=g0)
    pr = p.clone().requires_grad_(True)
    opt = {"rms_prop": lambda: torch.optim.RMSprop([pr], lr=5e-4),
           "adam": lambda: torch.optim.Adam([pr], lr=5e-4, betas=(0.9, 0.99)),
           "adamw": lambda: torch.optim.AdamW([pr], lr=5e-4, betas=(0.9, 0.99), weight_decay=0.01)}[name]()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step, norm, scratch = torch.zeros(4), torch.zeros(2), torch.zeros(1024)
    for it in range(5):
        g = torch.randn(n, generator=g0) * (10.0 ** (it - 3))
        pr.grad = g.clone()
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([pr], clip)
        opt.step()
        gk = g.clone()
        check(emu, emu.gg_optim_step(kind, p.data_ptr(), gk.data_ptr(), m.data_ptr(), v.data_ptr(), n, 5e-4, clip,
                                     step.data_ptr(), norm.data_ptr(), scratch.data_ptr(), None))
        if clip > 0:
            assert abs(norm[0].item() - g.norm().item()) <= 1e-5 * g.norm().item()
            torch.testing.assert_close(gk, pr.grad, rtol=1e-5, atol=1e-9)    # the clipped gradient is written back
        else:
            assert torch.equal(gk, g)                                         # untouched without clipping
        torch.testing.assert_close(p, pr.detach(), rtol=2e-5, atol=2e-7)
        assert step[0].item() == it + 1
    state = opt.state[pr]
    torch.testing.assert_close(v, state["square_avg"] if name == "rms_prop" else state["exp_avg_sq"], rtol=2e-5, atol=1e-12)
    if name != "rms_prop":
        torch.testing.assert_close(m, state["exp_avg"], rtol=2e-5, atol=1e-7)   # lerp of gradients up to 1e1: fp32 cancellation


def test_optimizer_rejects_misaligned_buffers_and_unknown_kind(emu):
    buf = torch.zeros(64)
    a = buf.data_ptr()
    assert emu.gg_optim_step(0, a + 4, a, a, a, 8, 1e-3, 0.0, a, None, None, None) == -1
    assert b"16-byte aligned" in emu.gg_last_error()
    assert emu.gg_optim_step(9, a, a, a, a, 8, 1e-3, 0.0, a, None, None, None) == -1
    assert emu.gg_optim_step(0, a, a, a, a, 8, 1e-3, 1.0, a, None, None, None) == -1      # clipping without scratch


class ShadowSeg(C.Structure):   # gemmgan_b200/csrc/kernels.h
    _fields_ = [("p_off", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32), ("col0", C.c_int32),
                ("ncols", C.c_int32), ("s_off", C.c_int64), ("s_ld", C.c_int64), ("transpose", C.c_int32),
                ("pad_", C.c_int32)]


def test_shadow_refresh_vector_and_scalar_paths(emu):
    """bf16 shadows of two parameter blocks of one flat buffer: an aligned [6, 16] block (vector path) and columns
    3..9 of a [5, 11] block (scalar path) into padded pitches; everything else in the shadow buffer stays untouched."""
    flat = torch.randn(6 * 16 + 5 * 11)
    shadow = torch.full((6 * 24 + 5 * 8,), 7.0, dtype=torch.bfloat16)
    segs = (ShadowSeg * 2)(ShadowSeg(0, 6, 16, 0, 16, 0, 24), ShadowSeg(96, 5, 11, 3, 7, 6 * 24, 8))
    bump = torch.zeros(1)
    check(emu, emu.emu_refresh_shadows(flat.data_ptr(), shadow.data_ptr(), C.addressof(segs), 2, bump.data_ptr()))
    a = shadow[: 6 * 24].view(6, 24)
    assert torch.equal(a[:, :16], flat[:96].view(6, 16).bfloat16()) and torch.all(a[:, 16:] == 7.0)
    b = shadow[6 * 24:].view(5, 8)
    assert torch.equal(b[:, :7], flat[96:].view(5, 11)[:, 3:10].bfloat16()) and torch.all(b[:, 7:] == 7.0)
    assert bump.item() == 1.0
    # transposed shadow (the K-major dgrad operands of the fused backward kernels): [5, 11] -> [11, 8-pitch]
    flat2 = torch.randn(5 * 11)
    sh2 = torch.full((11 * 8,), 7.0, dtype=torch.bfloat16)
    seg2 = (ShadowSeg * 1)(ShadowSeg(0, 5, 11, 0, 11, 0, 8, 1, 0))
    check(emu, emu.emu_refresh_shadows(flat2.data_ptr(), sh2.data_ptr(), C.addressof(seg2), 1, None))
    t = sh2.view(11, 8)
    assert torch.equal(t[:, :5], flat2.view(5, 11).t().bfloat16()) and torch.all(t[:, 5:] == 7.0)
