"""Thin torch-tensor wrappers over single C-ABI kernels (used by tests and micro-benchmarks)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Epilogue, GemmDesc


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(
    a, b, *, a_mn=False, b_mn=False, M=None, N=None, a2=None, b2=None,
    bias=None, pre=None, act=_lib.ACT_NONE, slope=0.0, drop_p=0.0, rng=None, site=0,
    mask=None, mask_pos=1.0, mask_neg=0.0, res=None, alpha=1.0,
    out_bf16=None, out_f32=None, accum=False, row_map=None,
    workspace=None, impl=_lib.IMPL_TCGEN05, splits=0, block_n=0, light=0, pair=0,
):
    """D = epilogue(alpha * (A @ B^T [+ A2 @ B2^T])).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16; b: [N,K] or [K,N] bf16 (2-D, last dim contiguous).
    """
    L = _lib.lib()
    d = GemmDesc()
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    d.M, d.N = M, N
    segs = [(a, b)] + ([(a2, b2)] if a2 is not None else [])
    d.nseg = len(segs)
    tf32 = a.dtype == torch.float32     # fp32 operands are multiplied as TF32 (gg_gemm_desc.tf32_operands)
    for i, (x, w) in enumerate(segs):
        assert x.dtype == w.dtype and x.dtype in (torch.bfloat16, torch.float32) and (x.dtype == torch.float32) == tf32
        assert x.stride(-1) == 1 and w.stride(-1) == 1
        K = x.shape[0] if a_mn else x.shape[1]
        Kb = w.shape[0] if b_mn else w.shape[1]
        assert K == Kb, (x.shape, w.shape)
        d.seg[i].a, d.seg[i].b = x.data_ptr(), w.data_ptr()
        d.seg[i].lda, d.seg[i].ldb = x.stride(0), w.stride(0)
        d.seg[i].K = K
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    e = d.epi
    e.alpha = alpha
    e.bias = _ptr(bias)
    if pre is not None:
        e.pre, e.pre_ld, e.pre_f32 = pre.data_ptr(), pre.stride(0), int(pre.dtype == torch.float32)
    e.act, e.slope = act, slope
    e.drop_p, e.rng, e.site = drop_p, _ptr(rng), site
    if mask is not None:
        e.mask, e.mask_ld, e.mask_f32 = mask.data_ptr(), mask.stride(0), int(mask.dtype == torch.float32)
    e.mask_pos, e.mask_neg = mask_pos, mask_neg
    if res is not None:
        e.res, e.res_ld, e.res_f32 = res.data_ptr(), res.stride(0), int(res.dtype == torch.float32)
    if out_bf16 is not None:
        e.out_bf16, e.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if out_f32 is not None:
        e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    e.accum_f32 = int(accum)
    if row_map is not None:
        e.row_div, e.row_mul, e.row_add = row_map
    if workspace is not None:
        d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    d.impl, d.force_splits, d.block_n, d.light, d.pair = impl, splits, block_n, light, pair
    d.tf32_operands = int(tf32)
    _lib.check(L.gg_gemm_bf16(C.byref(d), _stream()))


def wgrad_group(problems, workspace=None):
    """problems: list of (dy [K, M] bf16, x [K, N] bf16, out [M, N] fp32[, bias [M] fp32]). One launch (gg_wgrad_group)."""
    from . import _abi_decl as A

    L = _lib.lib()
    n = len(problems)
    items = (A.WgradItem * n)()
    total = 0
    for i, prob in enumerate(problems):
        dy, x, out = prob[:3]
        bias = prob[3] if len(prob) > 3 else None
        assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and out.dtype == torch.float32
        assert dy.shape[0] == x.shape[0] and out.shape == (dy.shape[1], x.shape[1])
        it = items[i]
        it.dy, it.ld_dy, it.x, it.ld_x = dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0)
        it.M, it.N, it.K = dy.shape[1], x.shape[1], dy.shape[0]
        it.out, it.ld = out.data_ptr(), out.stride(0)
        it.bias = None if bias is None else bias.data_ptr()
        total += it.M * it.N
    if workspace is None:
        workspace = torch.zeros(L.gg_wgrad_group_workspace_bytes(total), device=problems[0][0].device, dtype=torch.uint8)
    _lib.check(L.gg_wgrad_group(items, n, _ptr(workspace), workspace.numel(), _stream()))
    return workspace


def colsum_group(problems, workspace=None):
    """problems: list of (x [rows, N] bf16, out [N] fp32). One launch (gg_colsum_group)."""
    from . import _abi_decl as A

    L = _lib.lib()
    n = len(problems)
    items = (A.ColsumItem * n)()
    total = 0
    for i, (x, out) in enumerate(problems):
        assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and out.numel() == x.shape[1]
        it = items[i]
        it.inp, it.ld, it.rows, it.N, it.out = x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr()
        total += x.shape[1]
    if workspace is None:
        workspace = torch.zeros(L.gg_colsum_group_workspace_bytes(total), device=problems[0][0].device, dtype=torch.uint8)
    _lib.check(L.gg_colsum_group(items, n, _ptr(workspace), workspace.numel(), _stream()))
    return workspace


def xw_f32(x0, x1, w_bf16, workspace_mb=96):
    """[x0; x1] . w^T for two fp32 [B, K] matrices read in place and a bf16 [256, K] weight (gg_xw_f32) -> fp32 [2B, 256]."""
    L = _lib.lib()
    B, K = x0.shape
    assert x1.shape == x0.shape and w_bf16.shape == (256, K) and x0.is_contiguous() and x1.is_contiguous()
    assert w_bf16.stride(1) == 1 and w_bf16.stride(0) % 8 == 0
    out = torch.empty(2 * B, 256, device=x0.device, dtype=torch.float32)
    ws = torch.empty(workspace_mb << 20, device=x0.device, dtype=torch.uint8) if workspace_mb else None
    _lib.check(L.gg_xw_f32(x0.data_ptr(), x1.data_ptr(), B, K, w_bf16.data_ptr(), w_bf16.stride(0), out.data_ptr(),
                           None if ws is None else ws.data_ptr(), 0 if ws is None else ws.numel(), _stream()))
    return out


def dropout_bits(rng, site, p, n_elems):
    """Keep bits of a dropout site drawn once (gg_dropout_bits): uint32 words, bit idx = element idx."""
    L = _lib.lib()
    out = torch.empty(int(L.gg_dropout_bits_words(n_elems)), dtype=torch.int32, device=rng.device)
    _lib.check(L.gg_dropout_bits(rng.data_ptr(), site, float(p), n_elems, out.data_ptr(), _stream()))
    return out


def attention(qkv, nb, H, Lq, Lk=None, mask=None, drop_p=0.0, rng=None, site=0, dout=None, precomputed_bits=False):
    """Self-attention on a packed [nb*L, 3*H*hd] bf16 qkv tensor (the encoder-layer layout). Returns o, or
    (o, dqkv) when `dout` is given (gg_attention_fwd / gg_attention_bwd). precomputed_bits: draw the dropout mask
    once with gg_dropout_bits and hand it to both passes (what the engine does for 17 .. 320 tokens)."""
    from . import _abi_decl as A

    L = _lib.lib()
    Lk = Lk or Lq
    E = qkv.shape[1] // 3
    a = A.AttnArgs()
    a.q, a.ldq, a.q_mod = qkv.data_ptr(), qkv.stride(0), nb
    a.k, a.v, a.ldkv, a.kv_mod = qkv.data_ptr() + 2 * E, qkv.data_ptr() + 4 * E, qkv.stride(0), nb
    if mask is not None:
        a.mask, a.mask_mod = mask.data_ptr(), mask.shape[0]
    a.nb, a.H, a.hd, a.Lq, a.Lk = nb, H, E // H, Lq, Lk
    a.drop_p, a.rng, a.site = drop_p, (rng.data_ptr() if rng is not None else None), site
    if precomputed_bits and drop_p > 0.0:
        bits = dropout_bits(rng, site, drop_p, nb * H * Lq * Lk)
        a.dbits = bits.data_ptr()
    o = torch.empty(nb * Lq, E, device=qkv.device, dtype=torch.bfloat16)
    a.o, a.ldo = o.data_ptr(), E
    _lib.check(L.gg_attention_fwd(C.byref(a), _stream()))
    if dout is None:
        return o
    dqkv = torch.empty_like(qkv)
    stat = torch.empty(2 * nb * H * Lq, device=qkv.device, dtype=torch.float32)
    a.dout, a.lddo = dout.data_ptr(), dout.stride(0)
    a.dq, a.lddq = dqkv.data_ptr(), dqkv.stride(0)
    a.dk, a.dv, a.lddkv = dqkv.data_ptr() + 2 * E, dqkv.data_ptr() + 4 * E, dqkv.stride(0)
    a.stat = stat.data_ptr()
    _lib.check(L.gg_attention_bwd(C.byref(a), _stream()))
    return o, dqkv
