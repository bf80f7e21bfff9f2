"""GPU tests of the attention core (gg_attention_fwd / gg_attention_bwd) against torch SDPA + autograd in fp32 on
the same bf16 inputs, for the code paths: register-resident short self-attention (S <= 16, hd = 64: the tower at
8 patches), the tensor-core mid-size self-attention (17 <= S <= 128, hd = 64: the tower at 64 patches), the
flash-style long self-attention (129 <= S <= 320, hd = 64: the tower at 256 patches; 17 .. 19 tiles run with one warp
per tile), the generic
short-sequence path, and the shared-memory path for long sequences."""
import pytest
import torch

from gemmgan_b200 import ops

pytestmark = pytest.mark.gpu


def _reference(qkv, nb, H, L, mask, dout):
    E = qkv.shape[1] // 3
    hd = E // H
    x = qkv.float().requires_grad_(True)
    q, k, v = (x[:, i * E:(i + 1) * E].view(nb, L, H, hd).transpose(1, 2) for i in range(3))
    am = None
    if mask is not None:
        am = torch.zeros(nb, 1, 1, L, device=qkv.device)
        am.masked_fill_(mask.bool().view(nb, 1, 1, L), float("-inf"))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=am)
    o = o.transpose(1, 2).reshape(nb * L, E)
    o.backward(dout.float())
    return o.detach(), x.grad


@pytest.mark.parametrize("nb,H,hd,L", [(37, 4, 64, 9), (129, 4, 64, 16), (5, 4, 64, 3), (21, 4, 8, 6), (3, 4, 64, 65),
                                       (2, 2, 32, 257), (7, 4, 64, 17), (5, 4, 64, 32), (3, 4, 64, 48), (9, 4, 64, 80),
                                       (3, 4, 64, 96), (2, 4, 64, 128), (2, 4, 64, 129), (300, 4, 64, 65), (3, 4, 64, 257),
                                       (2, 4, 64, 200), (1, 4, 64, 320), (40, 4, 64, 257), (2, 4, 64, 280),
                                       (3, 4, 64, 300), (2, 4, 64, 272), (2, 4, 64, 305)])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_matches_sdpa(nb, H, hd, L, masked):
    g = torch.Generator(device="cuda").manual_seed(nb * 1000 + L)
    E = H * hd
    qkv = torch.randn(nb * L, 3 * E, device="cuda", generator=g).to(torch.bfloat16)
    dout = torch.randn(nb * L, E, device="cuda", generator=g).to(torch.bfloat16)
    mask = None
    if masked:
        mask = (torch.rand(nb, L, device="cuda", generator=g) < 0.3).to(torch.uint8)
        mask[:, 0] = 0  # the CLS key is never padded
    o, dqkv = ops.attention(qkv, nb, H, L, mask=mask, dout=dout)
    torch.cuda.synchronize()
    o_ref, g_ref = _reference(qkv, nb, H, L, mask, dout)
    assert (o.float() - o_ref).abs().max().item() <= 1e-2 * o_ref.abs().max().item()
    assert (dqkv.float() - g_ref).abs().max().item() <= 1.5e-2 * g_ref.abs().max().item()


@pytest.mark.parametrize("nb,L", [(64, 9), (16, 65), (8, 100), (4, 257), (3, 290)])
def test_attention_dropout_forward_backward_consistent(nb, L):
    """With dropout the backward must regenerate the forward's mask: d(sum o * w)/dv checked by linearity in v."""
    H, hd = 4, 64
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(nb * L, 3 * E, device="cuda", generator=g).to(torch.bfloat16)
    dout = torch.randn(nb * L, E, device="cuda", generator=g).to(torch.bfloat16)
    rng = torch.tensor([77, 5], device="cuda", dtype=torch.int64)
    o1, dqkv = ops.attention(qkv, nb, H, L, drop_p=0.3, rng=rng, site=2, dout=dout)
    o2 = ops.attention(qkv, nb, H, L, drop_p=0.3, rng=rng, site=2)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    # o is linear in v: <dout, o(v + t*dv)> - <dout, o(v)> = t * <grad_v, dv>
    dv = torch.randn(nb * L, E, device="cuda", generator=g).to(torch.bfloat16)
    qkv2 = qkv.clone()
    qkv2[:, 2 * E:] = (qkv[:, 2 * E:].float() + dv.float()).to(torch.bfloat16)
    dv_eff = qkv2[:, 2 * E:].float() - qkv[:, 2 * E:].float()
    o3 = ops.attention(qkv2, nb, H, L, drop_p=0.3, rng=rng, site=2)
    lhs = ((o3.float() - o1.float()) * dout.float()).sum().item()
    rhs = (dqkv[:, 2 * E:].float() * dv_eff).sum().item()
    assert abs(lhs - rhs) <= 3e-2 * abs(rhs) + 1.0
    # roughly 30 % of the probabilities are dropped: outputs differ from the dropout-free ones
    o0 = ops.attention(qkv, nb, H, L)
    assert (o0.float() - o1.float()).abs().mean().item() > 1e-3


@pytest.mark.parametrize("nb,L,masked", [(16, 65, False), (8, 100, True), (5, 17, False), (4, 257, True), (3, 290, False),
                                         (6, 129, True)])
def test_precomputed_dropout_bits_equal_the_in_kernel_masks(nb, L, masked):
    """gg_dropout_bits draws the attention-probability keep bits once per layer pass; the mid / long kernels read 16-bit
    windows of them in forward and backward (engine.cu::tower_forward). Same stream, same decisions: outputs and all three
    gradients are bitwise those of the kernels drawing Philox groups themselves."""
    H, hd = 4, 64
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(nb * L, 3 * E, device="cuda", generator=g).to(torch.bfloat16)
    dout = torch.randn(nb * L, E, device="cuda", generator=g).to(torch.bfloat16)
    mask = None
    if masked:
        mask = (torch.arange(L, device="cuda")[None, :] >= (L - torch.arange(nb, device="cuda") % 7)[:, None]).to(torch.uint8)
    rng = torch.tensor([1234567, 42], device="cuda", dtype=torch.int64)
    o1, d1 = ops.attention(qkv, nb, H, L, mask=mask, drop_p=0.1, rng=rng, site=8, dout=dout)
    o2, d2 = ops.attention(qkv, nb, H, L, mask=mask, drop_p=0.1, rng=rng, site=8, dout=dout, precomputed_bits=True)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and torch.equal(d1, d2)
    # the bit stream itself against the scalar definition (philox.cuh::dropout_keep) through a numpy port
    from test_gpu_enc_layer import philox_keep
    n = nb * H * L * L
    bits = ops.dropout_bits(rng, 8, 0.1, n).cpu().numpy().view("uint32")
    idx = torch.randint(0, n, (4096,), generator=torch.Generator().manual_seed(0)).numpy().astype("uint64")
    got = (bits[idx >> 5] >> (idx & 31).astype("uint32")) & 1
    want = philox_keep(1234567, 42, 8, idx, 0.1)
    assert (got.astype(bool) == want).all()


def _cross(q, kv, nb, H, Lk, mask=None, dout=None, kv_rows_shared=False):
    """Single-query cross-attention through the C ABI: q [nb, E], kv [nb(or nb/2)*Lk, 2E] (k | v)."""
    import ctypes as C

    from gemmgan_b200 import _abi_decl as A
    from gemmgan_b200 import _lib

    L = _lib.lib()
    E = q.shape[1]
    a = A.AttnArgs()
    a.q, a.ldq, a.q_mod = q.data_ptr(), q.stride(0), nb
    a.k, a.v, a.ldkv = kv.data_ptr(), kv.data_ptr() + 2 * E, kv.stride(0)
    a.kv_mod = kv.shape[0] // Lk
    if mask is not None:
        a.mask, a.mask_mod = mask.data_ptr(), mask.shape[0]
    a.nb, a.H, a.hd, a.Lq, a.Lk = nb, H, E // H, 1, Lk
    o = torch.empty(nb, E, device=q.device, dtype=torch.bfloat16)
    a.o, a.ldo = o.data_ptr(), E
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.gg_attention_fwd(C.byref(a), st))
    if dout is None:
        return o
    dq = torch.empty_like(q)
    dkv = torch.empty(nb * Lk, 2 * E, device=q.device, dtype=torch.bfloat16)   # per-row gradients (replicas unsummed)
    stat = torch.empty(2 * nb * H, device=q.device, dtype=torch.float32)
    a.dout, a.lddo = dout.data_ptr(), dout.stride(0)
    a.dq, a.lddq = dq.data_ptr(), dq.stride(0)
    a.dk, a.dv, a.lddkv = dkv.data_ptr(), dkv.data_ptr() + 2 * E, dkv.stride(0)
    a.stat = stat.data_ptr()
    _lib.check(L.gg_attention_bwd(C.byref(a), st))
    return o, dq, dkv


@pytest.mark.parametrize("nb,Lk", [(33, 17), (40, 32), (24, 65), (10, 100), (6, 128), (12, 9), (4, 200)])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shared_kv", [False, True])
def test_single_query_cross_attention(nb, Lk, masked, shared_kv):
    """patch2text / text2patch attention (Lq = 1): the warp-per-head kernel (17 <= Lk <= 128) and its neighbours;
    shared_kv: two replicas of the rows read the same keys / values (kv_mod = nb / 2), gradients stay per row."""
    H, hd = 4, 64
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(nb * 100 + Lk)
    nb -= nb % 2
    nkv = nb // 2 if shared_kv else nb
    q = torch.randn(nb, E, device="cuda", generator=g).to(torch.bfloat16)
    kv = torch.randn(nkv * Lk, 2 * E, device="cuda", generator=g).to(torch.bfloat16)
    dout = torch.randn(nb, E, device="cuda", generator=g).to(torch.bfloat16)
    mask = None
    if masked:
        mask = (torch.rand(nkv, Lk, device="cuda", generator=g) < 0.3).to(torch.uint8)
        mask[:, 0] = 0
    o, dq, dkv = _cross(q, kv, nb, H, Lk, mask=mask, dout=dout)
    torch.cuda.synchronize()
    qf = q.float().requires_grad_(True)
    kvf = kv.float().view(nkv, Lk, 2 * E)
    kvf = (kvf.repeat(2, 1, 1) if shared_kv else kvf).clone().requires_grad_(True)     # row b reads kv[b % nkv]
    qh = qf.view(nb, 1, H, hd).transpose(1, 2)
    kh = kvf[:, :, :E].reshape(nb, Lk, H, hd).transpose(1, 2)
    vh = kvf[:, :, E:].reshape(nb, Lk, H, hd).transpose(1, 2)
    am = None
    if mask is not None:
        mm = mask.bool().repeat(2, 1) if shared_kv else mask.bool()
        am = torch.zeros(nb, 1, 1, Lk, device="cuda").masked_fill_(mm.view(nb, 1, 1, Lk), float("-inf"))
    oref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=am).transpose(1, 2).reshape(nb, E)
    oref.backward(dout.float())
    assert (o.float() - oref).abs().max().item() <= 1e-2 * oref.abs().max().item()
    assert (dq.float() - qf.grad).abs().max().item() <= 1.5e-2 * qf.grad.abs().max().item()
    gref = kvf.grad.reshape(nb * Lk, 2 * E)
    assert (dkv.float() - gref).abs().max().item() <= 1.5e-2 * gref.abs().max().item()
