"""GPU tests of gg_gemm_bf16 (tcgen05 / TMEM / TMA kernel) through the C ABI: every epilogue feature, both
output paths (TMA-store boxes and direct per-row stores), ragged M / N / K, both operand majors, split-K,
the persistent multi-tile loop, and all three tile widths — against a torch fp32 reference on the same bf16
operands (tolerance: fp32 accumulation-order noise, plus one bf16 rounding step on bf16 outputs) and against
the CUDA-core check kernel that shares the epilogue code (dropout masks must agree bit for bit)."""
import pytest
import torch

from gemmgan_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def _mk(rows, cols, g, ld=None, scale=1.0):
    ld = ld or (cols + 7) // 8 * 8
    buf = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
    buf[:, :cols] = (torch.randn(rows, cols, device="cuda", generator=g) * scale).to(torch.bfloat16)
    return buf[:, :cols]


def _ref(a, b, a_mn, b_mn):
    af = a.float().t() if a_mn else a.float()
    bf = b.float().t() if b_mn else b.float()
    return af @ bf.t()


def _leaky(x, s):
    return torch.where(x > 0, x, s * x)


CASES = [
    # M, N, K, a_mn, b_mn, bn, splits
    (1000, 200, 256, 0, 0, 0, 0),
    (128, 64, 64, 0, 0, 64, 0),
    (300, 520, 136, 0, 1, 128, 0),
    (260, 1000, 200, 1, 1, 256, 0),
    (520, 264, 1000, 1, 0, 128, 3),
    (27648, 768, 256, 0, 0, 128, 0),     # persistent loop: ~9 tiles per CTA
    (18432, 512, 256, 0, 1, 256, 0),
    (256, 18872, 512, 1, 1, 256, 0),
]


PAIR_CASES = [
    # the CTA-pair configuration (tcgen05 cta_group::2, 256 x 256 tiles): ragged M / N / K tails, both operand
    # majors, split-K, the persistent multi-tile loop
    (1000, 300, 256, 0, 0, 256, 0),
    (256, 256, 64, 0, 0, 256, 0),
    (300, 520, 136, 0, 1, 256, 0),
    (260, 1000, 200, 1, 1, 256, 0),
    (520, 264, 1000, 1, 0, 256, 3),
    (27648, 768, 256, 0, 0, 256, 0),
    (18432, 512, 256, 0, 1, 256, 0),
    (256, 18872, 512, 1, 1, 256, 0),
    (2048, 256, 18872, 0, 0, 256, 9),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn,splits,pair", [c + (-1,) for c in CASES] + [c + (1,) for c in PAIR_CASES])
@pytest.mark.parametrize("aligned", [True, False])
def test_gemm_epilogue_features(M, N, K, a_mn, b_mn, bn, splits, aligned, pair):
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    a = _mk(K, M, g) if a_mn else _mk(M, K, g)
    b = _mk(K, N, g) if b_mn else _mk(N, K, g)
    bias = torch.randn(N, device="cuda", generator=g)
    pre = _mk(M, N, g)
    mask = _mk(M, N, g)
    res = torch.randn(M, N, device="cuda", generator=g)           # fp32 residual
    # aligned: pitches that allow TMA stores; unaligned: odd pitches force the direct-store path
    ld_b = (N + 7) // 8 * 8 + 8 if aligned else (N + 7) // 8 * 8 + 2
    ld_f = (N + 3) // 4 * 4 + 4 if aligned else (N + 3) // 4 * 4 + 1
    ob = torch.full((M, ld_b), 7.0, device="cuda", dtype=torch.bfloat16)
    of = torch.full((M, ld_f), 7.0, device="cuda", dtype=torch.float32)
    ws = torch.empty(max(1, splits) * M * N * 4 + 1024, device="cuda", dtype=torch.uint8)
    ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, pre=pre, act=_lib.ACT_LEAKY, slope=0.2,
             mask=mask, mask_pos=1.5, mask_neg=-0.5, res=res, alpha=0.5, out_bf16=ob[:, :N], out_f32=of[:, :N],
             workspace=ws, splits=splits, block_n=bn, pair=pair)
    torch.cuda.synchronize()
    want = _leaky(0.5 * _ref(a, b, a_mn, b_mn) + bias + pre.float(), 0.2)
    want = want * torch.where(mask.float() > 0, 1.5, -0.5) + res
    scale = want.abs().max().item()
    assert (of[:, :N] - want).abs().max().item() <= 1e-4 * scale
    assert (ob[:, :N].float() - want).abs().max().item() <= 5e-3 * scale
    # nothing outside the [M, N] window was touched (clipped TMA boxes, guarded direct stores). TMA stores clip
    # at 16-byte granularity: with an aligned pitch the row padding up to the next multiple of 8 bf16 (4 fp32)
    # columns may receive the zero accumulator tail (documented in include/gemmgan.h)
    nb = (N + 7) // 8 * 8 if aligned else N
    nf = (N + 3) // 4 * 4 if aligned else N
    if ld_b > nb:
        assert (ob[:, nb:] == 7.0).all()
    if ld_f > nf:
        assert (of[:, nf:] == 7.0).all()


def test_gemm_two_segments_row_map_and_film():
    g = torch.Generator(device="cuda").manual_seed(5)
    Bn, P, S, E, K, K2 = 37, 5, 6, 64, 192, 72
    M = Bn * P
    a, b = _mk(M, K, g), _mk(E, K, g)
    a2, b2 = _mk(M, K2, g), _mk(E, K2, g)
    out = torch.zeros(Bn * S, E, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, a2=a2, b2=b2, out_bf16=out, row_map=(P, S, 1))
    out2 = torch.zeros_like(out)
    ops.gemm(a, b, a2=a2, b2=b2, out_bf16=out2, row_map=(P, S, 1), pair=1)   # same through a CTA pair
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    want = (_ref(a, b, 0, 0) + _ref(a2, b2, 0, 0)).view(Bn, P, E)
    got = out.view(Bn, S, E)
    assert (got[:, 0] == 0).all()
    assert (got[:, 1:].float() - want).abs().max().item() <= 5e-3 * want.abs().max().item()
    # FiLM activation: tanh on the first half of the columns, clamp(-5, 5) on the second
    w = _mk(128, K, g, scale=2.0)
    of = torch.empty(M, 128, device="cuda", dtype=torch.float32)
    ops.gemm(a, w, act=_lib.ACT_FILM, out_f32=of)
    torch.cuda.synchronize()
    r = _ref(a, w, 0, 0)
    want = torch.cat([torch.tanh(r[:, :64]), r[:, 64:].clamp(-5, 5)], 1)
    assert (of - want).abs().max().item() <= 1e-4


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(27648, 768, 256, 0, 0), (18432, 256, 512, 0, 1), (1000, 200, 136, 0, 0),
                                              (300, 264, 64, 1, 1)])
def test_gemm_light_configuration(M, N, K, a_mn, b_mn):
    """The two-CTAs-per-SM configuration (4 epilogue warps, 2-stage ring) computes the same thing."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _mk(K, M, g) if a_mn else _mk(M, K, g)
    b = _mk(K, N, g) if b_mn else _mk(N, K, g)
    bias = torch.randn(N, device="cuda", generator=g)
    res = _mk(M, N, g)
    outs = []
    for light in (1, -1):
        ob = torch.zeros(M, (N + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
        of = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, act=_lib.ACT_LEAKY, slope=0.1, res=res,
                 out_bf16=ob[:, :N], block_n=128, light=light)
        ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, act=_lib.ACT_LEAKY, slope=0.1, res=res,
                 out_f32=of, block_n=128, light=light)
        outs.append((ob, of))
    torch.cuda.synchronize()
    want = _leaky(_ref(a, b, a_mn, b_mn) + bias, 0.1) + res.float()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert (outs[0][1] - want).abs().max().item() <= 1e-4 * want.abs().max().item()


@pytest.mark.parametrize("bn", [64, 128, 256, -256])
def test_gemm_dropout_matches_check_kernel(bn):
    """Dropout is a pure function of (seed, step, site, element index): the tensor-core kernel and the
    CUDA-core check kernel must drop exactly the same elements, whatever the tiling."""
    g = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 1111, 512, 128
    a, b = _mk(M, K, g), _mk(N, K, g)
    rng = torch.tensor([1234, 7], device="cuda", dtype=torch.int64)
    o1 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    o2 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, drop_p=0.25, rng=rng, site=3, out_f32=o1, block_n=abs(bn), pair=1 if bn < 0 else -1)  # -256: CTA pair
    ops.gemm(a, b, drop_p=0.25, rng=rng, site=3, out_f32=o2, impl=_lib.IMPL_SIMT_F32)
    torch.cuda.synchronize()
    assert ((o1 == 0) == (o2 == 0)).all()
    frac = (o1 == 0).float().mean().item()
    assert abs(frac - 0.25) < 0.01
    assert (o1 - o2).abs().max().item() <= 1e-3 * o2.abs().max().item()
    keep = o1 != 0
    want = _ref(a, b, 0, 0) / 0.75
    assert (o1[keep] - want[keep]).abs().max().item() <= 1e-3 * want.abs().max().item()


def test_wgrad_group_matches_reference_and_is_deterministic():
    """All weight gradients of one backward pass in one launch (ragged shapes, K-splits, TMA and direct stores)."""
    g = torch.Generator(device="cuda").manual_seed(21)
    shapes = [  # K (rows), M (out), N (in), out pitch
        (18432, 256, 512, 512), (18432, 512, 256, 256), (18432, 768, 256, 256), (9216, 256, 256, 256),
        (1024, 2048, 768, 768), (2048, 256, 256, 19124), (1000, 200, 72, 75), (64, 128, 128, 128), (300, 40, 1000, 1000),
    ]
    probs, refs, brefs = [], [], []
    for i, (K, M, N, ld) in enumerate(shapes):
        dy, x = _mk(K, M, g, scale=0.5), _mk(K, N, g, scale=0.5)
        buf = torch.full((M, ld), 3.0, device="cuda", dtype=torch.float32)
        # every other problem also asks for its bias gradient (column sums of dY, fused as a ones-column MMA)
        bias = torch.full((M + 3,), 9.0, device="cuda", dtype=torch.float32) if i % 2 == 0 else None
        probs.append((dy, x, buf[:, :N]) + ((bias[:M],) if bias is not None else ()))
        refs.append(dy.float().t() @ x.float())
        brefs.append(dy.double().sum(0))
    ws = ops.wgrad_group(probs)
    torch.cuda.synchronize()
    first = [p[2].clone() for p in probs]
    first_b = [p[3].clone() if len(p) > 3 else None for p in probs]
    for p, ref, bref in zip(probs, refs, brefs):
        dy, x, out = p[:3]
        assert (out - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-3, (dy.shape, x.shape)
        buf = out._base if out._base is not None else out
        if buf.shape[1] > out.shape[1]:
            assert (buf[:, out.shape[1]:] == 3.0).all()
        if len(p) > 3:
            assert (p[3].double() - bref).abs().max().item() <= 1e-5 * dy.shape[0] ** 0.5 * 4 + 1e-3, dy.shape
            assert (p[3]._base[dy.shape[1]:] == 9.0).all()      # nothing written behind the M bias elements
    # replay with the same workspace (counters must have been left at zero) -> bit-identical results
    for _ in range(3):
        for p in probs:
            p[2].fill_(-1.0)
            if len(p) > 3:
                p[3].fill_(-1.0)
        ops.wgrad_group(probs, workspace=ws)
        torch.cuda.synchronize()
        for p, f, fb in zip(probs, first, first_b):
            assert torch.equal(p[2], f)
            if fb is not None:
                assert torch.equal(p[3], fb)


def test_colsum_group_matches_reference_and_is_deterministic():
    g = torch.Generator(device="cuda").manual_seed(22)
    shapes = [(18432, 256), (18432, 768), (27648, 512), (1024, 2048), (1000, 203), (37, 64), (5000, 8), (2049, 100)]
    probs, refs = [], []
    for rows, N in shapes:
        x = _mk(rows, N, g)
        probs.append((x, torch.full((N,), 5.0, device="cuda", dtype=torch.float32)))
        refs.append(x.double().sum(0))
    ws = ops.colsum_group(probs)
    torch.cuda.synchronize()
    first = [p[1].clone() for p in probs]
    for (x, out), ref in zip(probs, refs):
        assert (out.double() - ref).abs().max().item() <= 1e-5 * x.shape[0] ** 0.5 * 4 + 1e-3
    for _ in range(3):
        for p in probs:
            p[1].fill_(-1.0)
        ops.colsum_group(probs, workspace=ws)
        torch.cuda.synchronize()
        for p, f in zip(probs, first):
            assert torch.equal(p[1], f)


@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 1000, 0), (2048, 256, 18868, 256), (300, 130, 97 * 4, 128),
                                      (128, 64, 36, 64)])
def test_gemm_with_fp32_operands_runs_as_tf32(M, N, K, bn):
    """gg_gemm_desc.tf32_operands: fp32 A / B read in place (no bf16 copy), multiplied on the tensor cores as TF32
    (10-bit mantissa), fp32 accumulate. Against fp64 torch: relative error of the TF32 rounding of the operands."""
    import torch
    from gemmgan_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(64 << 20, device="cuda", dtype=torch.uint8)
    ops.gemm(a, b, bias=bias, out_f32=out, block_n=bn, workspace=ws)
    torch.cuda.synchronize()
    want = (a.double() @ b.double().t() + bias.double()).float()
    err = (out - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-3, err                       # (bf16 operands give ~1e-2 here)
