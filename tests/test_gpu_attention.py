"""GPU tests of the attention core (gg_attention_fwd / gg_attention_bwd) against torch SDPA + autograd in fp32 on
the same bf16 inputs, for the code paths: register-resident short self-attention (S <= 16, hd = 64: the tower at
8 patches), the tensor-core mid-size self-attention (17 <= S <= 128, hd = 64: the tower at 64 patches), the generic
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
                                       (3, 4, 64, 96), (2, 4, 64, 128), (2, 4, 64, 129), (300, 4, 64, 65)])
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


@pytest.mark.parametrize("nb,L", [(64, 9), (16, 65), (8, 100)])
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
