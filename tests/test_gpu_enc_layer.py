"""The fused encoder-layer forward kernel (gemmgan_b200/csrc/enc_layer.cu, gg_encoder_layer_fwd) on a B200 against
plain PyTorch fp32 math of nn.TransformerEncoderLayer's post-norm branch (reference:
src/conditional_gan_cross_attention_with_film.py:114-119, :144) on the same bf16-rounded operands — every tensor the
kernel writes (output and the ones saved for the backward), with key-padding masks, with / without biases, partial
save ranges, tail tiles, and with dropout ON: the kernel's Philox masks are regenerated on the host (numpy port of
csrc/philox.cuh) and applied in the torch reference, so the comparison stays exact in the dropout configuration the
training step runs.

Tolerances: bf16 outputs of fp32-accumulated bf16 products -> 1e-2 of the tensor's max (one bf16 ulp is 4e-3)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gemmgan_b200 import _abi_decl as A
from gemmgan_b200 import _lib

pytestmark = pytest.mark.gpu
E, F, NH, HD = 256, 512, 4, 64


def philox_keep(seed, step, site, idx, p):
    """keep decision of element `idx` (numpy int64 array) at a dropout site: csrc/philox.cuh dropout_keep."""
    idx = np.asarray(idx, dtype=np.uint64)
    grp = idx >> np.uint64(3)
    c = [grp & np.uint64(0xFFFFFFFF), grp >> np.uint64(32), np.full_like(grp, site), np.full_like(grp, step & 0xFFFFFFFF)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(7):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    sel = (idx & np.uint64(7)).astype(np.int64)
    words = np.stack(c, axis=-1)
    w = np.take_along_axis(words, (sel >> 1)[..., None], axis=-1)[..., 0]
    u = np.where(sel & 1, w >> np.uint64(16), w & np.uint64(0xFFFF))
    return u >= np.uint64(int(p * 65536.0 + 0.5))


def reference(x, W, mask, nb, S, p, seed, step, site):
    """fp32 torch math on bf16-rounded operands; returns every tensor the kernel writes."""
    dev = x.device
    xf = x.float()
    rows = nb * S
    keep_scale = 1.0 / (1.0 - p) if p > 0 else 1.0

    def keep(site_k, shape, width):
        if p == 0:
            return torch.ones(shape, device=dev)
        idx = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
        return torch.from_numpy(philox_keep(seed, step, site_k, idx, p)).to(dev).float() * keep_scale

    qkv = xf @ W["w_in"].float().t() + (W["b_in"] if W["b_in"] is not None else 0)
    qkv_b = qkv.bfloat16().float()                      # the attention reads the stored bf16 q / k / v
    q, k, v = (t.view(nb, S, NH, HD).transpose(1, 2) for t in qkv_b.split(E, dim=1))
    s = (q @ k.transpose(-1, -2)) * 0.125
    if mask is not None:
        mm = mask[(torch.arange(nb, device=dev) % mask.shape[0])]
        s = s.masked_fill(mm[:, None, None, :].bool(), float("-inf"))
    pr = torch.softmax(s, dim=-1)
    pr = (pr * keep(site, (nb, NH, S, S), S)).bfloat16().float()
    ao = (pr @ v).transpose(1, 2).reshape(rows, E)
    ao_b = ao.bfloat16().float()
    sa = ao_b @ W["w_out"].float().t() + (W["b_out"] if W["b_out"] is not None else 0)
    z1 = xf + sa * keep(site + 1, (rows, E), E)
    mu1, var1 = z1.mean(1, keepdim=True), z1.var(1, unbiased=False, keepdim=True)
    rs1 = torch.rsqrt(var1 + 1e-5)
    x1 = (z1.bfloat16().float() - mu1) * rs1 * W["g1"] + (W["be1"] if W["be1"] is not None else 0)
    x1_b = x1.bfloat16().float()
    h = torch.relu(x1_b @ W["w_ff1"].float().t() + (W["b_ff1"] if W["b_ff1"] is not None else 0))
    h = h * keep(site + 2, (rows, F), F)
    h_b = h.bfloat16().float()
    ff = h_b @ W["w_ff2"].float().t() + (W["b_ff2"] if W["b_ff2"] is not None else 0)
    z2 = x1_b + ff * keep(site + 3, (rows, E), E)
    mu2, var2 = z2.mean(1, keepdim=True), z2.var(1, unbiased=False, keepdim=True)
    rs2 = torch.rsqrt(var2 + 1e-5)
    out = (z2.bfloat16().float() - mu2) * rs2 * W["g2"] + (W["be2"] if W["be2"] is not None else 0)
    return dict(qkv=qkv, ao=ao, z1=z1, x1=x1, h=h, z2=z2, out=out, mean1=mu1[:, 0], rstd1=rs1[:, 0], mean2=mu2[:, 0],
                rstd2=rs2[:, 0])


def make_weights(bias, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale
    W = dict(w_in=r(3 * E, E, scale=E ** -0.5).bfloat16(), w_out=r(E, E, scale=E ** -0.5).bfloat16(),
             w_ff1=r(F, E, scale=E ** -0.5).bfloat16(), w_ff2=r(E, F, scale=F ** -0.5).bfloat16(),
             g1=1 + 0.1 * r(E), g2=1 + 0.1 * r(E))
    for k, n in (("b_in", 3 * E), ("b_out", E), ("b_ff1", F), ("b_ff2", E), ("be1", E), ("be2", E)):
        W[k] = 0.1 * r(n) if bias else None
    return W


def run_kernel(x, W, mask, nb, S, p, seed, step, site, save_rows, precomputed_bits=False):
    L = _lib.lib()
    rows = nb * S
    dev = x.device
    out = {k: torch.full((rows, w), float("nan"), device=dev, dtype=torch.bfloat16)
           for k, w in (("qkv", 3 * E), ("ao", E), ("z1", E), ("x1", E), ("h", F), ("z2", E), ("out", E))}
    for k in ("mean1", "rstd1", "mean2", "rstd2"):
        out[k] = torch.full((rows,), float("nan"), device=dev, dtype=torch.float32)
    rng = torch.tensor([seed, step], dtype=torch.int64, device=dev)
    P = A.EncLayerParams()
    P.nb, P.S, P.E, P.F, P.n_heads, P.save_rows = nb, S, E, F, NH, save_rows
    P.x = x.data_ptr()
    for k in ("w_in", "w_out", "w_ff1", "w_ff2"):
        setattr(P, k, W[k].data_ptr())
        setattr(P, "ld_" + k[2:], W[k].shape[1])
    for k in ("b_in", "b_out", "b_ff1", "b_ff2", "g1", "be1", "g2", "be2"):
        setattr(P, k, None if W[k] is None else W[k].data_ptr())
    P.mask = None if mask is None else mask.data_ptr()
    P.mask_mod = 0 if mask is None else mask.shape[0]
    P.drop_p, P.eps, P.rng, P.site = p, 1e-5, rng.data_ptr(), site
    for k in out:
        setattr(P, k, out[k].data_ptr())
    if precomputed_bits:   # the engine's way: the three per-element masks drawn once by gg_dropout_bits
        from gemmgan_b200 import ops
        keep = [ops.dropout_bits(rng, site + 1, p, rows * E), ops.dropout_bits(rng, site + 2, p, rows * F),
                ops.dropout_bits(rng, site + 3, p, rows * E)]
        P.dbits1, P.dbits2, P.dbits3 = (k.data_ptr() for k in keep)
    _lib.check(L.gg_encoder_layer_fwd(C.byref(P), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-12)


@pytest.mark.parametrize("nb,S,bias,masked,p,save", [
    (40, 9, True, True, 0.0, -1),        # three tiles of 14 / 14 / 12 sequences
    (3072, 9, True, False, 0.0, -1),     # the critic tower of BASELINE config 3: 3 * 1024 sequences, 220 tiles
    (33, 9, False, True, 0.0, -1),       # bias=False towers (film / cross / img variants)
    (40, 9, True, True, 0.1, -1),        # dropout on, masks regenerated on the host
    (300, 9, True, True, 0.1, 9 * 200),  # only the first 200 sequences are saved for the backward
    (29, 9, True, True, 0.0, 0),         # output only (generator tower inside a critic step)
    (37, 16, True, True, 0.1, -1),       # 16 tokens: 8 sequences per tile, no unused rows
    (50, 5, True, True, 0.0, -1),        # 25 sequences per tile (125 of 128 rows)
    (7, 1, True, False, 0.0, -1),        # degenerate: one token per sequence
])
def test_fused_layer_matches_torch(nb, S, bias, masked, p, save):
    _lib.require_device(0)
    torch.manual_seed(1)
    rows = nb * S
    x = torch.randn(rows, E, device="cuda").bfloat16()
    W = make_weights(bias)
    mask = None
    if masked and S > 1:
        nm = max(1, nb // 3)                                    # replicas share masks: sequence b uses row b % nm
        k = torch.randint(0, S, (nm,), device="cuda")
        mask = (torch.arange(S, device="cuda")[None, :] >= (S - k)[:, None]).to(torch.uint8).contiguous()
        mask[:, 0] = 0                                          # the CLS key is never padded
    seed, step, site = 0x1234ABCD5678, 7, 24
    got = run_kernel(x, W, mask, nb, S, p, seed, step, site, save)
    want = reference(x, W, mask, nb, S, p, seed, step, site)
    n_save = rows if save < 0 else min(save, rows)
    assert rel(got["out"], want["out"]) < 1.5e-2, rel(got["out"], want["out"])
    for k in ("qkv", "ao", "z1", "x1", "h", "z2"):
        if n_save:
            assert rel(got[k][:n_save], want[k][:n_save]) < 1.5e-2, (k, rel(got[k][:n_save], want[k][:n_save]))
        assert torch.isnan(got[k][n_save:].float()).all(), k     # rows beyond the save range are not written
    for k in ("mean1", "rstd1", "mean2", "rstd2"):
        if n_save:
            assert torch.allclose(got[k][:n_save], want[k][:n_save], rtol=2e-2, atol=2e-3), k
        assert torch.isnan(got[k][n_save:]).all(), k


@pytest.mark.parametrize("nb,S,save", [(40, 9, -1), (300, 9, 9 * 200), (37, 16, -1), (3072, 9, 0)])
def test_fused_layer_with_precomputed_dropout_bits_is_bitwise_the_same(nb, S, save):
    """Reading the keep bits gg_dropout_bits drew (what the engine does: the masks depend only on the step counter and are
    drawn next to the tower head) instead of running Philox in the epilogues: every output tensor bit for bit."""
    _lib.require_device(0)
    torch.manual_seed(2)
    x = torch.randn(nb * S, E, device="cuda").bfloat16()
    W = make_weights(True)
    a = run_kernel(x, W, None, nb, S, 0.1, 0x77AA, 3, 16, save)
    b = run_kernel(x, W, None, nb, S, 0.1, 0x77AA, 3, 16, save, precomputed_bits=True)
    for k in a:
        assert torch.equal(torch.nan_to_num(a[k].float(), nan=-7.0), torch.nan_to_num(b[k].float(), nan=-7.0)), k


def test_fused_layer_is_deterministic_and_dropout_changes_with_the_step():
    _lib.require_device(0)
    torch.manual_seed(2)
    nb, S = 100, 9
    x = torch.randn(nb * S, E, device="cuda").bfloat16()
    W = make_weights(True)
    a = run_kernel(x, W, None, nb, S, 0.1, 11, 3, 8, -1)
    b = run_kernel(x, W, None, nb, S, 0.1, 11, 3, 8, -1)
    c = run_kernel(x, W, None, nb, S, 0.1, 11, 4, 8, -1)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert not torch.equal(a["out"], c["out"])


# ------------------------------------------------------------------ fused ffn-half backward (gg_encoder_ffn_bwd)
@pytest.mark.parametrize("rows,p", [(1000, 0.0), (18432, 0.1), (333, 0.1), (128, 0.0)])
def test_fused_ffn_backward_matches_torch(rows, p):
    """gz = LN2'(dout), gy = mask(gz), gh = (gy W2) * [h > 0] / (1 - p), gb = gz + gh W1 against fp32 torch math on the
    same bf16 operands (dropout mask of the residual branch regenerated on the host; the ffn dropout acts through the
    stored activation h, whose zeros are dropped or relu-inactive units)."""
    _lib.require_device(0)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    r = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale
    dout, z2 = r(rows, E).bfloat16(), r(rows, E).bfloat16()
    gamma = 1 + 0.1 * r(E)
    w1, w2 = r(F, E, scale=E ** -0.5).bfloat16(), r(E, F, scale=F ** -0.5).bfloat16()
    h = torch.relu(r(rows, F))
    h = (h * (torch.rand(rows, F, device="cuda", generator=g) > 0.1)).bfloat16()      # zeros = inactive or dropped
    zf = z2.float()
    mean, var = zf.mean(1), zf.var(1, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    seed, step, site = 0xABCDEF12345, 5, 19
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    gh = torch.full((rows, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    gb = torch.full((rows, E), float("nan"), device="cuda", dtype=torch.bfloat16)
    rng = torch.tensor([seed, step], dtype=torch.int64, device="cuda")
    P = A.EncFfnBwdParams()
    P.rows, P.dout, P.z2, P.mean2, P.rstd2, P.gamma2 = rows, dout.data_ptr(), z2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
    P.h, P.w2t, P.ld_w2t, P.w1t, P.ld_w1t = h.data_ptr(), w2t.data_ptr(), E, w1t.data_ptr(), F
    P.drop_p, P.rng, P.site, P.gh, P.gb = p, rng.data_ptr(), site, gh.data_ptr(), gb.data_ptr()
    _lib.check(L.gg_encoder_ffn_bwd(C.byref(P), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # reference
    d = dout.float() * gamma
    xh = (zf - mean[:, None]) * rstd[:, None]
    gz = rstd[:, None] * (d - d.mean(1, keepdim=True) - xh * (d * xh).mean(1, keepdim=True))
    keep_scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    if p > 0:
        idx = np.arange(rows * E, dtype=np.int64).reshape(rows, E)
        keep = torch.from_numpy(philox_keep(seed, step, site, idx, p)).cuda().float() * keep_scale
    else:
        keep = torch.ones(rows, E, device="cuda")
    gy = (gz * keep).bfloat16().float()
    gh_ref = (gy @ w2.float()) * (h.float() > 0) * keep_scale
    gb_ref = gz + gh_ref.bfloat16().float() @ w1.float()
    assert rel(gh, gh_ref) < 1.5e-2, rel(gh, gh_ref)
    assert rel(gb, gb_ref) < 1.5e-2, rel(gb, gb_ref)
