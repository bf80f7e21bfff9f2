"""The residual-add + dropout + LayerNorm kernels of the encoder layers (gemmgan_b200/csrc/layernorm.cu; the four
`x = LN(x + Drop(sublayer(x)))` of nn.TransformerEncoderLayer per layer and pass, reference
src/conditional_gan_cross_attention_with_film.py:114-119) checked WITHOUT a GPU: the unchanged .cu is
compiled for the host (tests/cuda_emu/emu.h) and compared with torch's layer_norm forward / autograd backward in fp32.
Tolerances are those of bf16 storage (z, out, dz are bf16 tensors): 1e-2 of the tensor's scale; fp32 statistics and
parameter gradients 2e-3. On the B200 the same kernels are covered through the step parity tests."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import emu_build


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("layernorm", tmp_path_factory.mktemp("cuda_emu"))
    vp, i32, i64, f32, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32
    L.emu_add_ln_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, f32, vp, u32]
    L.emu_add_ln_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, vp, u32, vp]
    L.emu_ln_bwd_scratch_floats.argtypes = [i64, i32]
    L.emu_ln_bwd_scratch_floats.restype = i64
    return L


def ptr(t):
    return None if t is None else t.data_ptr()


def run(emu, x, y, w, b, dout, drop_p=0.0, rng=None, site=3):
    rows, E = x.shape
    z, out, dz, dy = (torch.empty_like(x) for _ in range(4))
    mean, rstd = torch.empty(rows), torch.empty(rows)
    dw, db = torch.empty(E), torch.empty(E)
    scratch = torch.empty(int(emu.emu_ln_bwd_scratch_floats(rows, E)))
    rc = emu.emu_add_ln_fwd(ptr(x), ptr(y), ptr(w), ptr(b), ptr(z), ptr(out), ptr(mean), ptr(rstd), rows, E, 1e-5,
                            drop_p, ptr(rng), site)
    assert rc == 0, emu.gg_last_error()
    rc = emu.emu_add_ln_bwd(ptr(dout), ptr(z), ptr(mean), ptr(rstd), ptr(w), ptr(dz), ptr(dy) if drop_p > 0 else None,
                            ptr(dw), ptr(db), rows, E, drop_p, ptr(rng), site, ptr(scratch))
    assert rc == 0, emu.gg_last_error()
    return dict(z=z, out=out, mean=mean, rstd=rstd, dz=dz, dy=dy, dw=dw, db=db)


def close(got, want, tol):
    scale = want.abs().max().item() + 1e-12
    err = (got.float() - want.float()).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("E", [32, 96, 256, 512, 1024])
@pytest.mark.parametrize("rows", [1, 7, 300])
def test_add_layernorm_forward_backward_match_torch(emu, rows, E):
    g = torch.Generator().manual_seed(rows * 7 + E)
    x = torch.randn(rows, E, generator=g).bfloat16()
    y = (0.5 * torch.randn(rows, E, generator=g)).bfloat16()
    w = 1.0 + 0.1 * torch.randn(E, generator=g)
    b = 0.1 * torch.randn(E, generator=g)
    dout = torch.randn(rows, E, generator=g).bfloat16()
    r = run(emu, x, y, w, b, dout)
    zf = x.float() + y.float()
    close(r["z"], zf, 4e-3)                                   # bf16 rounding of the stored sum
    zr = r["z"].float().requires_grad_(True)                  # torch on the kernel's own (rounded) z ...
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = F.layer_norm(zf, (E,), w, b, 1e-5)
    close(r["out"], out, 1e-2)
    close(r["mean"], zf.mean(1), 2e-3)
    close(r["rstd"], 1.0 / torch.sqrt(zf.var(1, unbiased=False) + 1e-5), 2e-3)
    # ... for the backward, which recomputes the normalised rows from the saved z / mean / rstd
    ref = F.layer_norm(zr, (E,), wr, br, 1e-5)
    (ref * dout.float()).sum().backward()
    close(r["dz"], zr.grad, 1.5e-2)
    close(r["dw"], wr.grad, 5e-3)
    close(r["db"], br.grad, 2e-3)


def test_dropout_mask_is_shared_by_forward_and_backward(emu):
    rows, E, p = 64, 256, 0.25
    g = torch.Generator().manual_seed(5)
    x = torch.randn(rows, E, generator=g).bfloat16()
    y = (torch.randn(rows, E, generator=g).abs() + 0.5).bfloat16()     # never zero: a kept y always changes z
    w, b = torch.ones(E), torch.zeros(E)
    dout = torch.randn(rows, E, generator=g).bfloat16()
    rng = torch.tensor([1234, 7], dtype=torch.int64)                   # (seed, step) of the engine's dropout stream
    a = run(emu, x, y, w, b, dout, drop_p=p, rng=rng)
    again = run(emu, x, y, w, b, dout, drop_p=p, rng=rng)
    assert torch.equal(a["z"], again["z"]) and torch.equal(a["dy"], again["dy"])       # deterministic in (seed, step)
    kept = a["z"] != x
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02, frac
    # kept entries carry y / (1 - p); the backward passes dz / (1 - p) through exactly the same entries
    close(a["z"][kept], (x.float() + y.float() / (1 - p))[kept], 8e-3)
    assert torch.all(a["dy"][~kept] == 0)
    close(a["dy"][kept], a["dz"].float()[kept] / (1 - p), 8e-3)
    other = run(emu, x, y, w, b, dout, drop_p=p, rng=torch.tensor([1234, 8], dtype=torch.int64))
    assert not torch.equal(other["z"], a["z"])                          # a new step draws a new mask


def test_unsupported_width_is_an_error(emu):
    t = torch.zeros(4, 48).bfloat16()
    f = torch.zeros(48)
    assert emu.emu_add_ln_fwd(ptr(t), ptr(t), ptr(f), ptr(f), ptr(t), ptr(t), ptr(f), ptr(f), 4, 48, 1e-5, 0.0, None,
                              0) == -1
    assert b"unsupported" in emu.gg_last_error()
