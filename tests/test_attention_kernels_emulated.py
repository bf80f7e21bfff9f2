"""The attention kernels of the hot path (gemmgan_b200/csrc/attention.cu; SURVEY.md §8 a7 / a8: masked self-attention
of the encoder layers and the single-query patch2text / text2patch attention,
src/conditional_gan_cross_attention_with_film.py:114-123, :144-152) checked WITHOUT a GPU. The unchanged kernels are
compiled for the host (tests/cuda_emu/emu.h); their five PTX wrappers (cp.async, ldmatrix, ldmatrix.trans, mma.sync
m16n8k16 bf16, ex2) are replaced by host versions that follow the PTX fragment layouts (tests/cuda_emu/emu_attention.cpp),
so the register-resident short kernel, the mma.sync mid kernel, the flash-style long kernel with its recomputed
backward, the single-query kernels and the generic paths all run thread for thread and are compared with torch SDPA +
autograd in fp32 on the same bf16 inputs — the comparison tests/test_gpu_attention.py makes on the B200, at sizes the
emulation finishes in seconds. Tolerances as there: 1e-2 / 1.5e-2 of the tensor's scale (bf16 outputs)."""
import ctypes as C

import pytest
import torch

import emu_build
from gemmgan_b200 import _abi_decl as A


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("attention", tmp_path_factory.mktemp("cuda_emu"))
    L.gg_attention_fwd.argtypes = [C.POINTER(A.AttnArgs), C.c_void_p]
    L.gg_attention_bwd.argtypes = [C.POINTER(A.AttnArgs), C.c_void_p]
    return L


def check(L, rc):
    assert rc == 0, L.gg_last_error()


def self_attention(L, qkv, nb, H, S, mask=None, drop_p=0.0, rng=None, site=0, dout=None, precomputed_bits=False):
    """gemmgan_b200.ops.attention on host tensors: packed [nb*S, 3*H*hd] bf16 qkv (the encoder-layer layout)."""
    E = qkv.shape[1] // 3
    a = A.AttnArgs()
    a.q, a.ldq, a.q_mod = qkv.data_ptr(), qkv.stride(0), nb
    a.k, a.v, a.ldkv, a.kv_mod = qkv.data_ptr() + 2 * E, qkv.data_ptr() + 4 * E, qkv.stride(0), nb
    if mask is not None:
        a.mask, a.mask_mod = mask.data_ptr(), mask.shape[0]
    a.nb, a.H, a.hd, a.Lq, a.Lk = nb, H, E // H, S, S
    a.drop_p, a.rng, a.site = drop_p, (rng.data_ptr() if rng is not None else None), site
    if precomputed_bits:   # the keep bits of the site drawn once (gg_dropout_bits), read by forward and backward
        L.gg_dropout_bits_words.restype = C.c_int64
        L.gg_dropout_bits_words.argtypes = [C.c_int64]
        L.gg_dropout_bits.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
        n = nb * H * S * S
        bits = torch.empty(int(L.gg_dropout_bits_words(n)), dtype=torch.int32)
        check(L, L.gg_dropout_bits(rng.data_ptr(), site, drop_p, n, bits.data_ptr(), None))
        a.dbits = bits.data_ptr()
    o = torch.empty(nb * S, E, dtype=torch.bfloat16)
    a.o, a.ldo = o.data_ptr(), E
    check(L, L.gg_attention_fwd(C.byref(a), None))
    if dout is None:
        return o
    dqkv = torch.empty_like(qkv)
    stat = torch.empty(2 * nb * H * S)
    a.dout, a.lddo = dout.data_ptr(), dout.stride(0)
    a.dq, a.lddq = dqkv.data_ptr(), dqkv.stride(0)
    a.dk, a.dv, a.lddkv = dqkv.data_ptr() + 2 * E, dqkv.data_ptr() + 4 * E, dqkv.stride(0)
    a.stat = stat.data_ptr()
    check(L, L.gg_attention_bwd(C.byref(a), None))
    return o, dqkv


def reference(qkv, nb, H, S, mask, dout):
    E = qkv.shape[1] // 3
    hd = E // H
    x = qkv.float().requires_grad_(True)
    q, k, v = (x[:, i * E:(i + 1) * E].view(nb, S, H, hd).transpose(1, 2) for i in range(3))
    am = None
    if mask is not None:
        am = torch.zeros(nb, 1, 1, S).masked_fill_(mask.bool().view(nb, 1, 1, S), float("-inf"))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=am)
    o = o.transpose(1, 2).reshape(nb * S, E)
    o.backward(dout.float())
    return o.detach(), x.grad


# short register kernel (S <= 16), mma.sync mid kernel (17..128), flash-style long kernel (129..320; 257 = the
# reference's 256 patches + CLS runs with 17 warps, 288 / 304 with 18 / 19, 305 and 320 = 20 tiles on 16 warps),
# generic paths (hd != 64)
@pytest.mark.parametrize("nb,H,hd,S", [(5, 4, 64, 9), (3, 4, 64, 16), (2, 4, 64, 3), (3, 4, 8, 6), (2, 4, 64, 17),
                                       (2, 4, 64, 65), (1, 4, 64, 128), (1, 2, 64, 129), (1, 2, 64, 257),
                                       (1, 1, 64, 288), (1, 1, 64, 304), (1, 1, 64, 305), (1, 1, 64, 320),
                                       (1, 2, 32, 40)])
@pytest.mark.parametrize("masked", [False, True])
def test_self_attention_matches_sdpa(emu, nb, H, hd, S, masked):
    g = torch.Generator().manual_seed(nb * 1000 + S)
    E = H * hd
    qkv = torch.randn(nb * S, 3 * E, generator=g).bfloat16()
    dout = torch.randn(nb * S, E, generator=g).bfloat16()
    mask = None
    if masked:
        mask = (torch.rand(nb, S, generator=g) < 0.3).to(torch.uint8)
        mask[:, 0] = 0                                   # the CLS key is never padded
    o, dqkv = self_attention(emu, qkv, nb, H, S, mask=mask, dout=dout)
    o_ref, g_ref = reference(qkv, nb, H, S, mask, dout)
    assert (o.float() - o_ref).abs().max().item() <= 1e-2 * o_ref.abs().max().item()
    assert (dqkv.float() - g_ref).abs().max().item() <= 1.5e-2 * g_ref.abs().max().item()


@pytest.mark.parametrize("nb,S,masked", [(3, 65, True), (2, 17, False), (2, 140, True), (1, 257, False)])
def test_precomputed_dropout_bits_equal_the_in_kernel_masks(emu, nb, S, masked):
    """The mid / long kernels reading 16-bit windows of the once-drawn keep bits (AttnArgs.dbits) make exactly the
    decisions they make when they draw the Philox groups themselves: outputs and gradients bitwise equal."""
    H, E = 4, 256
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(nb * S, 3 * E, generator=g).bfloat16()
    dout = torch.randn(nb * S, E, generator=g).bfloat16()
    mask = None
    if masked:
        mask = (torch.arange(S)[None, :] >= (S - 1 - torch.arange(nb) * 3)[:, None]).to(torch.uint8)
    rng = torch.tensor([99, 12], dtype=torch.int64)
    o1, d1 = self_attention(emu, qkv, nb, H, S, mask=mask, drop_p=0.1, rng=rng, site=8, dout=dout)
    o2, d2 = self_attention(emu, qkv, nb, H, S, mask=mask, drop_p=0.1, rng=rng, site=8, dout=dout, precomputed_bits=True)
    assert torch.equal(o1, o2) and torch.equal(d1, d2)


@pytest.mark.parametrize("nb,S", [(6, 9), (2, 65), (1, 140)])
def test_dropout_mask_is_regenerated_by_the_backward(emu, nb, S):
    """o is linear in v: <dout, o(v + dv)> - <dout, o(v)> = <grad_v, dv> holds only if the backward redraws the
    forward's dropout mask; same (seed, step, site) -> same output."""
    H, hd = 4, 64
    E = H * hd
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(nb * S, 3 * E, generator=g).bfloat16()
    dout = torch.randn(nb * S, E, generator=g).bfloat16()
    rng = torch.tensor([77, 5], dtype=torch.int64)
    o1, dqkv = self_attention(emu, qkv, nb, H, S, drop_p=0.3, rng=rng, site=2, dout=dout)
    assert torch.equal(o1, self_attention(emu, qkv, nb, H, S, drop_p=0.3, rng=rng, site=2))
    dv = torch.randn(nb * S, E, generator=g).bfloat16()
    qkv2 = qkv.clone()
    qkv2[:, 2 * E:] = (qkv[:, 2 * E:].float() + dv.float()).bfloat16()
    dv_eff = qkv2[:, 2 * E:].float() - qkv[:, 2 * E:].float()
    o3 = self_attention(emu, qkv2, nb, H, S, drop_p=0.3, rng=rng, site=2)
    lhs = ((o3.float() - o1.float()) * dout.float()).sum().item()
    rhs = (dqkv[:, 2 * E:].float() * dv_eff).sum().item()
    assert abs(lhs - rhs) <= 3e-2 * abs(rhs) + 1.0
    o0 = self_attention(emu, qkv, nb, H, S)
    assert (o0.float() - o1.float()).abs().mean().item() > 1e-3


@pytest.mark.parametrize("nb,Lk", [(6, 9), (8, 17), (4, 65), (2, 128), (2, 200)])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shared_kv", [False, True])
def test_single_query_cross_attention(emu, nb, Lk, masked, shared_kv):
    """patch2text / text2patch attention (Lq = 1, :149-152); shared_kv: two replicas of the rows read the same keys /
    values (kv_mod = nb / 2) and the gradients stay per row."""
    H, hd = 4, 64
    E = H * hd
    g = torch.Generator().manual_seed(nb * 100 + Lk)
    nkv = nb // 2 if shared_kv else nb
    q = torch.randn(nb, E, generator=g).bfloat16()
    kv = torch.randn(nkv * Lk, 2 * E, generator=g).bfloat16()
    dout = torch.randn(nb, E, generator=g).bfloat16()
    mask = None
    if masked:
        mask = (torch.rand(nkv, Lk, generator=g) < 0.3).to(torch.uint8)
        mask[:, 0] = 0
    a = A.AttnArgs()
    a.q, a.ldq, a.q_mod = q.data_ptr(), q.stride(0), nb
    a.k, a.v, a.ldkv, a.kv_mod = kv.data_ptr(), kv.data_ptr() + 2 * E, kv.stride(0), nkv
    if mask is not None:
        a.mask, a.mask_mod = mask.data_ptr(), mask.shape[0]
    a.nb, a.H, a.hd, a.Lq, a.Lk = nb, H, hd, 1, Lk
    o = torch.empty(nb, E, dtype=torch.bfloat16)
    a.o, a.ldo = o.data_ptr(), E
    check(emu, emu.gg_attention_fwd(C.byref(a), None))
    dq, dkv, stat = torch.empty_like(q), torch.empty(nb * Lk, 2 * E, dtype=torch.bfloat16), torch.empty(2 * nb * H)
    a.dout, a.lddo, a.dq, a.lddq = dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0)
    a.dk, a.dv, a.lddkv, a.stat = dkv.data_ptr(), dkv.data_ptr() + 2 * E, dkv.stride(0), stat.data_ptr()
    check(emu, emu.gg_attention_bwd(C.byref(a), None))
    qf = q.float().requires_grad_(True)
    kvf = kv.float().view(nkv, Lk, 2 * E)
    kvf = (kvf.repeat(2, 1, 1) if shared_kv else kvf).clone().requires_grad_(True)      # row b reads kv[b % nkv]
    qh = qf.view(nb, 1, H, hd).transpose(1, 2)
    kh = kvf[:, :, :E].reshape(nb, Lk, H, hd).transpose(1, 2)
    vh = kvf[:, :, E:].reshape(nb, Lk, H, hd).transpose(1, 2)
    am = None
    if mask is not None:
        mm = mask.bool().repeat(2, 1) if shared_kv else mask.bool()
        am = torch.zeros(nb, 1, 1, Lk).masked_fill_(mm.view(nb, 1, 1, Lk), float("-inf"))
    oref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=am).transpose(1, 2).reshape(nb, E)
    oref.backward(dout.float())
    assert (o.float() - oref).abs().max().item() <= 1e-2 * oref.abs().max().item()
    assert (dq.float() - qf.grad).abs().max().item() <= 1.5e-2 * qf.grad.abs().max().item()
    gref = kvf.grad.reshape(nb * Lk, 2 * E)
    assert (dkv.float() - gref).abs().max().item() <= 1.5e-2 * gref.abs().max().item()
