"""The HBM-bound glue kernels of the training step (gemmgan_b200/csrc/elementwise.cu) checked WITHOUT a GPU: the
unchanged .cu is compiled for the host (tests/cuda_emu/emu.h) and each kernel is compared with the torch expression of
the reference it replaces. The centre piece is the gradient-penalty chain: the
kernels that surround the GEMMs of the Gram-matrix formulation — trunk-1 combine (alpha mix after W1x), u2, the per-row
norm / penalty / r_b kernel, the loss reduction — are driven with torch matmuls standing in for the tensor-core GEMMs and
must reproduce `WGAN_GP.gradient_penalty` + the double backward of the reference (autograd,
src/vanilla_gan_unconditional.py:304-327, :381) on the same weights and inputs.

Tolerances: bf16 storage of activations (h1, h2, u2, ru1, dv1) -> 1e-2 of the tensor's scale; fp32 reductions 1e-5.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import emu_build

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGS = {
    "cast_f32_bf16": [vp, i64, vp, i64, i64, i32],
    "mask_with_cls": [vp, vp, i32, i32],
    "film_apply": [vp, vp, vp, i32, i32, i32],
    "film_bwd": [vp, vp, vp, vp, i32, i32, i32],
    "assemble_tokens": [vp, vp, i32, i32, i32, i32, vp],
    "unassemble_tokens": [vp, vp, i32, i32, i32, i32],
    "relu_bwd": [vp, vp, vp, i64],
    "sum_replicas": [vp, vp, i32, i64],
    "colsum": [vp, i32, i64, i64, i32, vp, f32, vp, i32, vp],
    "trunk1_combine": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32],
    "rowdot_bias": [vp, vp, vp, vp, i32, i32],
    "gp_u2": [vp, vp, vp, i32, i32, f32],
    "gp_rows": [vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, f32],
    "score_bwd": [vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, f32],
    "disc_losses": [vp, vp, vp, i32, f32, f32],
    "gen_loss": [vp, vp, i32, f32],
    "embed_gather": [vp, vp, vp, vp, i32, i32, vp, i32, i32],
    "embed_grad": [vp, vp, vp, i32, i32, vp, vp, i32, i32],
    "masked_mean_rows": [vp, vp, vp, i32, i32, i32],
    "gather_rows": [vp, i64, vp, vp, i64, i64, i32],
    "bn_fwd": [vp, i64, vp, vp, vp, vp, f32, f32, i32, vp, i64, vp, vp, i32, i32],
    "bn_bwd": [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, i32],
}


class Emu:
    def __init__(self, L):
        self.L = L
        for name, sig in SIGS.items():
            getattr(L, "emu_" + name).argtypes = sig

    def __getattr__(self, name):
        fn = getattr(self.L, "emu_" + name)

        def call(*args):
            rc = fn(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args])
            assert rc == 0, self.L.gg_last_error()
        return call


@pytest.fixture(scope="module")
def k(tmp_path_factory):
    return Emu(emu_build.build("elementwise", tmp_path_factory.mktemp("cuda_emu")))


def close(got, want, tol):
    scale = want.abs().max().item() + 1e-12
    err = (got.float() - want.float()).abs().max().item()
    assert err <= tol * scale, (err, scale)


def fro(got, want):
    return ((got.float() - want.float()).norm() / (want.float().norm() + 1e-12)).item()


def leaky(x, slope):
    return torch.where(x > 0, x, slope * x)


# ------------------------------------------------------------------------------------------ gradient penalty
@pytest.mark.parametrize("slope", [0.0, 0.2])
@pytest.mark.parametrize("B,G,H", [(8, 203, 32), (37, 1000, 256)])
def test_gradient_penalty_chain_matches_autograd(k, B, G, H, slope):
    g = torch.Generator().manual_seed(B + G)
    real, fake = torch.randn(B, G, generator=g), torch.randn(B, G, generator=g)
    alpha = torch.rand(B, 1, generator=g)
    W1 = torch.randn(H, G, generator=g) / G ** 0.5
    b1 = 0.1 * torch.randn(H, generator=g)
    W2 = torch.randn(H, H, generator=g) / H ** 0.5
    b2 = 0.1 * torch.randn(H, generator=g)
    w3 = torch.randn(H, generator=g) / H ** 0.5
    b3 = torch.tensor([0.3])
    gpw = 10.0

    # ---- the reference: gradient_penalty (:304-327) + what disc_loss.backward() (:381) makes of it, by autograd
    W1r, W2r, w3r = (t.clone().requires_grad_(True) for t in (W1, W2, w3))

    def critic(x):
        return leaky(leaky(x @ W1r.T + b1, slope) @ W2r.T + b2, slope) @ w3r + b3

    xhat = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    grad = torch.autograd.grad(critic(xhat).sum(), xhat, create_graph=True)[0]
    n_ref = grad.norm(2, dim=1)
    gp_ref = ((n_ref - 1) ** 2).mean()
    dW1_ref, dW2_ref, dw3_ref = torch.autograd.grad(gpw * gp_ref, (W1r, W2r, w3r))
    with torch.no_grad():
        d_fake, d_real = critic(fake), critic(real)

    # ---- the engine's sequence; torch matmuls stand in for the tcgen05 GEMMs
    a1x = torch.cat([fake, real]) @ W1.T                                    # one GEMM over [fake; real]
    h1 = torch.empty(3 * B, H, dtype=torch.bfloat16)
    k.trunk1_combine(a1x, None, b1, alpha.contiguous(), h1, B, H, 3, 1, slope)
    want_h1 = leaky(torch.cat([fake, real, xhat.detach()]) @ W1.T + b1, slope)
    close(h1, want_h1, 1e-2)
    h2f = leaky(h1.float() @ W2.T + b2, slope)
    score = torch.empty(3 * B)
    k.rowdot_bias(h2f, w3, b3, score, 3 * B, H)
    close(score[:B], d_fake, 2e-2)
    close(score[B:2 * B], d_real, 2e-2)
    h1i, h2i = h1[2 * B:].contiguous(), h2f[2 * B:].bfloat16().contiguous()
    u2 = torch.empty(B, H, dtype=torch.bfloat16)
    k.gp_u2(h2i, w3, u2, B, H, slope)
    m2 = torch.where(h2i.float() > 0, 1.0, slope)
    close(u2, m2 * w3, 1e-2)
    m1 = torch.where(h1i.float() > 0, 1.0, slope)
    u1 = m1 * (u2.float() @ W2)                                             # dgrad GEMM with the mask epilogue
    M = W1 @ W1.T                                                           # Gram matrix of W1x
    y = u1 @ M
    norms, pen = torch.empty(B), torch.empty(B)
    ru1, dv1 = torch.empty(B, H, dtype=torch.bfloat16), torch.empty(B, H, dtype=torch.bfloat16)
    k.gp_rows(y.contiguous(), u1.contiguous(), h1i, norms, pen, ru1, dv1, B, H, slope, gpw, 1.0 / B)
    close(norms, n_ref.detach(), 1.5e-2)                                    # ||dD/dx_hat|| per row, no [B, G] tensor
    close(pen, (n_ref.detach() - 1) ** 2, 4e-2)
    stats = torch.zeros(16)
    k.disc_losses(score, pen, stats, B, gpw, 1.0 / B)
    assert stats[2].item() == pytest.approx(gp_ref.item(), rel=3e-2)
    assert stats[0].item() == pytest.approx(-d_real.mean().item(), abs=2e-2)
    assert stats[1].item() == pytest.approx(d_fake.mean().item(), abs=2e-2)
    assert stats[3].item() == pytest.approx(stats[0].item() + stats[1].item() + gpw * stats[2].item(), rel=1e-6)
    # gradients of gp_weight * GP: dW1x = (U1^T diag(r) U1) W1x, and through u1 = m1 * (u2 W2):
    # dW2 = u2^T dv1, dw3 = sum_b m2 * (dv1 W2^T)
    r = gpw * 2.0 / B * (1 - 1 / norms)
    close(ru1, r[:, None] * u1, 1e-2)
    close(dv1, m1 * r[:, None] * y, 1e-2)
    Q = u1.T @ ru1.float()                                                  # [H, H], rides on the W1 wgrad GEMM
    # ReLU / LeakyReLU masks taken from bf16 activations flip for pre-activations within bf16 rounding of 0, which moves
    # single entries by O(1) of their size: gradients are compared by relative Frobenius error (as in test_gpu_parity)
    assert fro(Q @ W1, dW1_ref) < 0.08
    assert fro(u2.float().T @ dv1.float(), dW2_ref) < 0.08
    assert fro((m2 * (dv1.float() @ W2.T)).sum(0), dw3_ref) < 0.08


def test_score_backward_and_generator_loss(k):
    B, H, slope = 19, 64, 0.2
    g = torch.Generator().manual_seed(2)
    h2 = torch.randn(2 * B, H, generator=g).bfloat16()
    w3 = torch.randn(H, generator=g)
    da2, roww = torch.empty(2 * B, H, dtype=torch.bfloat16), torch.empty(2 * B)
    k.score_bwd(h2, w3, da2, roww, 2 * B, B, H, slope, 1.0, -1.0, 1.0 / B)      # d(mean D(fake) - mean D(real))/da2
    sign = torch.cat([torch.ones(B), -torch.ones(B)]) / B
    close(da2, sign[:, None] * w3 * torch.where(h2.float() > 0, 1.0, slope), 1e-2)
    assert torch.allclose(roww, sign)
    score, stats = torch.randn(300, generator=g), torch.zeros(16)
    k.gen_loss(score, stats, 300, 1.0 / 300)
    assert stats[4].item() == pytest.approx(-score.mean().item(), abs=1e-6)     # G_loss = -mean D(G(z)) (:34-36)


# ------------------------------------------------------------------------------------------ FiLM (:129-137)
@pytest.mark.parametrize("B,P,Dp", [(3, 5, 32), (4, 8, 1024), (1, 1, 8)])
def test_film_apply_and_backward(k, B, P, Dp):
    g = torch.Generator().manual_seed(Dp + P)
    patches = torch.randn(B, P, Dp, generator=g).bfloat16()
    gamma = torch.tanh(torch.randn(B, Dp, generator=g))
    beta = torch.clamp(4 * torch.randn(B, Dp, generator=g), -5, 5)              # some entries sit on the clamp
    gb = torch.cat([gamma, beta], 1).contiguous()
    mod = torch.empty_like(patches)
    k.film_apply(patches, gb, mod, B, P, Dp)
    close(mod, gamma[:, None] * patches.float() + beta[:, None], 8e-3)
    dmod = torch.randn(B, P, Dp, generator=g).bfloat16()
    dgb = torch.empty(B, 2 * Dp, dtype=torch.bfloat16)
    k.film_bwd(dmod, patches, gb, dgb, B, P, Dp)
    # gradients w.r.t. the PRE-activations of tanh / clamp
    want_g = (dmod.float() * patches.float()).sum(1) * (1 - gamma ** 2)
    want_b = dmod.float().sum(1) * (beta.abs() < 5)
    close(dgb[:, :Dp], want_g, 1e-2)
    close(dgb[:, Dp:], want_b, 1e-2)


# ------------------------------------------------------------------------------------------ token plumbing
@pytest.mark.parametrize("R", [1, 3])
def test_token_assembly_and_its_backward(k, R):
    B, S, E = 4, 6, 32
    g = torch.Generator().manual_seed(R)
    cls = torch.randn(E, generator=g)
    tokens = torch.randn(B, S - 1, E, generator=g).bfloat16()
    x = torch.full((R, B, S, E), 9.0, dtype=torch.bfloat16)
    x[0, :, 1:] = tokens                                            # the patch projection wrote replica 0 (:139-142)
    k.assemble_tokens(x, cls, R, B, S, E, None)
    want = torch.cat([cls.bfloat16().expand(B, 1, E), tokens], 1).expand(R, B, S, E)
    assert torch.equal(x, want)
    x2 = torch.full((R, B, S, E), 9.0, dtype=torch.bfloat16)
    k.assemble_tokens(x2, cls, R, B, S, E, tokens)                  # token rows of every replica from `src`
    assert torch.equal(x2, want)
    dx = torch.randn(R, B, S, E, generator=g).bfloat16()
    dpe = torch.empty(B, S - 1, E, dtype=torch.bfloat16)
    k.unassemble_tokens(dx, dpe, R, B, S, E)
    close(dpe, dx.float().sum(0)[:, 1:], 8e-3)
    pad = torch.rand(B, S - 1, generator=g) > 0.5
    out = torch.empty(B, S, dtype=torch.uint8)
    k.mask_with_cls(pad.to(torch.uint8), out, B, S - 1)
    assert torch.equal(out.bool(), torch.cat([torch.zeros(B, 1, dtype=torch.bool), pad], 1))   # CLS never padded


# ------------------------------------------------------------------------------------------ casts / sums / gathers
@pytest.mark.parametrize("rows,cols,ld_src,ld_dst", [(5, 16, 16, 16), (7, 203, 208, 208), (3, 8, 20, 12), (2, 7, 7, 9)])
def test_cast_with_pitches(k, rows, cols, ld_src, ld_dst):
    src = torch.randn(rows, ld_src)
    dst = torch.full((rows, ld_dst), 5.0, dtype=torch.bfloat16)
    k.cast_f32_bf16(src, ld_src, dst, ld_dst, rows, cols)
    assert torch.equal(dst[:, :cols], src[:, :cols].bfloat16()) and torch.all(dst[:, cols:] == 5.0)


@pytest.mark.parametrize("rows,N,f32_in", [(1, 1, 1), (300, 70, 0), (5000, 33, 1), (17, 256, 0)])
def test_column_sums(k, rows, N, f32_in):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, N + 3, generator=g)
    x = x if f32_in else x.bfloat16()
    roww = torch.rand(rows, generator=g)
    out, scratch = torch.full((N,), 2.0), torch.empty(64 * N)
    k.colsum(x, f32_in, N + 3, rows, N, roww, 0.5, out, 1, scratch)
    want = 2.0 + 0.5 * (roww[:, None] * x.float()[:, :N]).sum(0)
    close(out, want, 2e-5)
    k.colsum(x, f32_in, N + 3, rows, N, None, 1.0, out, 0, scratch)
    close(out, x.float()[:, :N].sum(0), 2e-5)


def test_replica_sum_relu_backward_masked_mean(k):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(3, 1001, generator=g).bfloat16()
    out = torch.empty(1001, dtype=torch.bfloat16)
    k.sum_replicas(a, out, 3, 1001)
    close(out, a.float().sum(0), 8e-3)
    grad, h = torch.randn(777, generator=g).bfloat16(), torch.randn(777, generator=g).bfloat16()
    res = torch.empty(777, dtype=torch.bfloat16)
    k.relu_bwd(grad, h, res, 777)
    assert torch.equal(res, torch.where(h.float() > 0, grad, torch.zeros_like(grad)))
    B, P, D = 5, 7, 130
    x = torch.randn(B, P, D, generator=g)
    pad = torch.arange(P)[None, :] >= torch.tensor([7, 1, 3, 6, 2])[:, None]    # True = padding, patch 0 always kept
    mean = torch.empty(B, D)
    k.masked_mean_rows(x, pad.to(torch.uint8), mean, B, P, D)
    want = (x * (~pad)[:, :, None]).sum(1) / (~pad).sum(1, keepdim=True)        # conditional_gan_concat.py:137-138
    close(mean, want, 1e-6)
    k.masked_mean_rows(x, None, mean, B, P, D)
    close(mean, x.mean(1), 1e-6)


def test_embedding_gather_and_deterministic_table_gradient(k):
    B, Eh, V0, V1 = 23, 128, 10, 7
    g = torch.Generator().manual_seed(4)
    e0 = torch.randn(V0, Eh, generator=g).requires_grad_(True)
    e1 = torch.randn(V1, Eh, generator=g).requires_grad_(True)
    y0, y1 = torch.randint(0, V0, (B,), generator=g), torch.randint(0, V1, (B,), generator=g)
    c = torch.empty(B, 2 * Eh, dtype=torch.bfloat16)
    k.embed_gather(e0.detach(), e1.detach(), y0, y1, V0, V1, c, B, Eh)
    ref = torch.cat([F.embedding(y0, e0), F.embedding(y1, e1)], 1)              # benchmark_generative_model.py:138-150
    assert torch.equal(c, ref.detach().bfloat16())
    dc = torch.randn(B, 2 * Eh, generator=g).bfloat16()
    (ref * dc.float()).sum().backward()
    g0, g1 = torch.empty(V0, Eh), torch.empty(V1, Eh)
    k.embed_grad(dc, y0, y1, V0, V1, g0, g1, B, Eh)
    close(g0, e0.grad, 1e-6)
    close(g1, e1.grad, 1e-6)


class ColsumItem(C.Structure):   # gg_colsum_item, include/gemmgan.h
    _fields_ = [("inp", vp), ("ld", i64), ("rows", i64), ("N", i32), ("out", vp)]


def test_grouped_column_sums_are_exact_and_deterministic(k):
    """Every bias gradient of one backward pass in one launch (gg_colsum_group): problems of different shapes, the
    16-byte and the scalar path, one chunk and several (last-arriver reduction in chunk order), arrival counters left
    at zero so that the next launch can reuse the workspace."""
    L = k.L
    L.emu_colsum_group.argtypes = [vp, i32, vp, i64]
    L.emu_colsum_group_workspace_bytes.argtypes = [i64]
    L.emu_colsum_group_workspace_bytes.restype = i64
    g = torch.Generator().manual_seed(9)
    shapes = [(300, 256, 256), (5000, 70, 72), (4100, 64, 67), (9, 1, 8), (2049, 130, 136)]   # rows, N, pitch
    xs = [torch.randn(r, ld, generator=g).bfloat16() for r, _, ld in shapes]
    outs = [torch.full((n,), -3.0) for _, n, _ in shapes]
    items = (ColsumItem * len(shapes))(*[ColsumItem(x.data_ptr(), ld, r, n, o.data_ptr())
                                         for x, o, (r, n, ld) in zip(xs, outs, shapes)])
    nbytes = int(L.emu_colsum_group_workspace_bytes(sum(n for _, n, _ in shapes)))
    ws = torch.zeros(nbytes, dtype=torch.uint8)
    assert L.emu_colsum_group(C.addressof(items), len(shapes), ws.data_ptr(), nbytes) == 0, L.gg_last_error()
    first = [o.clone() for o in outs]
    for x, o, (r, n, ld) in zip(xs, outs, shapes):
        close(o, x.float()[:, :n].sum(0), 2e-5)
    assert torch.all(ws[: 64 * 1024] == 0)                       # counters are back at zero
    for o in outs:
        o.fill_(7.0)
    assert L.emu_colsum_group(C.addressof(items), len(shapes), ws.data_ptr(), nbytes) == 0
    assert all(torch.equal(a, b) for a, b in zip(first, outs))   # bit-identical on the second launch
    assert L.emu_colsum_group(C.addressof(items), 41, ws.data_ptr(), nbytes) == -1   # more than COLSUM_GROUP_MAX


@pytest.mark.parametrize("cols,ld_src,ld_dst", [(1024, 1024, 1024), (203, 203, 211), (768, 772, 768)])
def test_gather_rows_builds_a_batch(k, cols, ld_src, ld_dst):
    """gg_gather_rows (device-side batch assembly): picked rows are copied, -1 rows are the zero
    padding of src/multi_patch_multi_token_gan_dataloader.py:37-38; vector and scalar paths, pitched tensors."""
    g = torch.Generator().manual_seed(0)
    n_src, rows = 57, 40
    src = torch.randn(n_src, ld_src, generator=g)
    index = torch.randint(-1, n_src, (rows,), generator=g, dtype=torch.int64)
    index[3] = -1
    dst = torch.full((rows, ld_dst), 7.0)
    k.gather_rows(src, ld_src, index, dst, ld_dst, rows, cols)
    want = torch.where((index >= 0)[:, None], src[index.clamp(min=0), :cols], torch.zeros(rows, cols))
    assert torch.equal(dst[:, :cols], want)
    assert torch.all(dst[:, cols:] == 7.0)            # the pitch padding is left alone


@pytest.mark.parametrize("B,E", [(8, 32), (37, 100), (64, 256)])
def test_batchnorm_forward_backward_and_running_statistics(k, B, E):
    """k_bn_fwd / k_bn_bwd against nn.BatchNorm1d (the generator's attn_bn of src/conditional_gan_attention.py:108, :126):
    training-mode output, in-place running_mean / running_var update (unbiased variance, momentum 0.1), eval-mode output
    from the running statistics, and the training-mode backward (dx, dweight, dbias) against autograd."""
    g = torch.Generator().manual_seed(B + E)
    x = (torch.randn(B, E, generator=g) * 2 + 0.5).bfloat16()
    bn = torch.nn.BatchNorm1d(E)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(E, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(E, generator=g))
        bn.running_mean.copy_(torch.randn(E, generator=g))
        bn.running_var.copy_(torch.rand(E, generator=g) + 0.5)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    xr = x.float().requires_grad_(True)
    want = bn(xr)
    y = torch.empty(B, E, dtype=torch.bfloat16)
    mean, rstd = torch.empty(E), torch.empty(E)
    k.bn_fwd(x, E, bn.weight.detach(), bn.bias.detach(), rm, rv, 0.1, 1e-5, 1, y, E, mean, rstd, B, E)
    assert torch.allclose(y.float(), want.detach(), atol=2e-2, rtol=1e-2)          # bf16 output
    assert torch.allclose(mean, x.float().mean(0), atol=1e-5)
    assert torch.allclose(rm, bn.running_mean, atol=1e-5) and torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    dy = torch.randn(B, E, generator=g).bfloat16()
    want.backward(dy.float())
    dx = torch.empty(B, E, dtype=torch.bfloat16)
    dgamma, dbeta = torch.empty(E), torch.empty(E)
    k.bn_bwd(dy, E, x, E, mean, rstd, bn.weight.detach(), dx, E, dgamma, dbeta, B, E)
    assert torch.allclose(dgamma, bn.weight.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(dbeta, bn.bias.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(dx.float(), xr.grad, atol=2e-2 * xr.grad.abs().max().item())
    # eval mode: the running statistics, untouched
    bn.eval()
    rm2, rv2 = rm.clone(), rv.clone()
    k.bn_fwd(x, E, bn.weight.detach(), bn.bias.detach(), rm2, rv2, 0.1, 1e-5, 0, y, E, mean, rstd, B, E)
    assert torch.allclose(y.float(), bn(x.float()).detach(), atol=2e-2, rtol=1e-2)
    assert torch.equal(rm2, rm) and torch.equal(rv2, rv)
