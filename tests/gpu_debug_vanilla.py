"""Diagnostic (not a pytest file): closed-form check of every intermediate of the vanilla critic step."""
import sys
import time

t0 = time.time()
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
print("import torch", time.time() - t0, flush=True)
from oracle import restated  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def main():
    cfg = T.SMALL if len(sys.argv) < 2 else getattr(T, sys.argv[1])
    slope = 0.0
    t1 = time.time()
    o, t = T.build_pair("vanilla", cfg, "adam", slope)
    print("build_pair", time.time() - t1, flush=True)
    B, G, L, H = cfg["B"], cfg["G"], cfg["latent"], cfg["hidden"]
    x, cond = restated.synthetic_batch("vanilla", B, G, seed=5)
    g = torch.Generator().manual_seed(99)
    z = torch.randn(B, L, generator=g)
    alpha = torch.rand(B, 1, generator=g)
    W1 = o.disc.discriminator[0][0].weight.detach().clone()
    b1 = o.disc.discriminator[0][0].bias.detach().clone()
    W2 = o.disc.discriminator[1][0].weight.detach().clone()
    b2 = o.disc.discriminator[1][0].bias.detach().clone()
    w3 = o.disc.final_layer.weight.detach().clone()[0]
    t1 = time.time()
    o.train_disc(x, z, cond, alpha)
    print("oracle step", time.time() - t1, flush=True)
    dev = t.device
    t1 = time.time()
    t.train_disc(x.to(dev), z.to(dev), alpha=alpha.to(dev))
    torch.cuda.synchronize()
    print("gpu step", time.time() - t1, flush=True)
    eng = t._engine(B)
    fake = o.last["fake"]
    # closed form (fp32, CPU)
    xi = alpha * x + (1 - alpha) * fake
    def fwd(inp):
        a1 = inp @ W1.t() + b1
        h1 = torch.where(a1 > 0, a1, slope * a1)
        a2 = h1 @ W2.t() + b2
        h2 = torch.where(a2 > 0, a2, slope * a2)
        return a1, h1, a2, h2
    a1i, h1i, a2i, h2i = fwd(xi)
    m1 = torch.where(a1i > 0, 1.0, slope)
    m2 = torch.where(a2i > 0, 1.0, slope)
    u2 = m2 * w3
    u1 = m1 * (u2 @ W2)
    M = W1 @ W1.t()
    y = u1 @ M
    n = (y * u1).sum(1).sqrt()
    r = 10.0 * (2.0 / B) * (1 - 1 / n)
    dv1 = m1 * (r[:, None] * y)
    Q = (r[:, None] * u1).t() @ u1
    a1f, h1f, a2f, h2f = fwd(fake)
    a1r, h1r, a2r, h2r = fwd(x)
    da2 = torch.cat([(1.0 / B) * w3 * torch.where(a2f > 0, 1.0, slope), (-1.0 / B) * w3 * torch.where(a2r > 0, 1.0, slope)])
    h1fr = torch.cat([h1f, h1r])
    da1 = (da2 @ W2) * torch.where(torch.cat([a1f, a1r]) > 0, 1.0, slope)
    dW1_loss = da1.t() @ torch.cat([fake, x])
    dW1_gp = Q @ W1
    print("gram", rel(eng.buffer("gram"), M))
    print("u2", rel(eng.buffer("u2"), u2))
    print("u1f", rel(eng.buffer("u1f"), u1))
    print("y", rel(eng.buffer("y"), y))
    print("norms", rel(eng.buffer("gp_norms")[:, 0], n))
    print("dv1", rel(eng.buffer("dv1"), dv1))
    print("ru1", rel(eng.buffer("ru1"), r[:, None] * u1))
    print("Qb", rel(eng.buffer("Qb"), Q))
    print("da2", rel(eng.buffer("da2"), da2))
    print("da1", rel(eng.buffer("da1"), da1))
    # mask-flip hypothesis: redo the checks with the ENGINE's own activation masks / inputs
    h1e = eng.buffer("h1").float().cpu()
    da2e = eng.buffer("da2").float().cpu()
    da1_chk = (da2e @ W2.bfloat16().float()) * torch.where(h1e[:2 * B] > 0, 1.0, slope)
    print("da1 vs own-mask closed form", rel(eng.buffer("da1"), da1_chk))
    flips = ((h1e[:2 * B] > 0) != (torch.cat([a1f, a1r]) > 0)).sum().item()
    print("mask flips in h1(fake,real):", flips, "of", h1e[:2 * B].numel())
    xe = torch.cat([eng.buffer("fake_bf16").float().cpu(), eng.buffer("real_bf16").float().cpu()])
    dW1_chk = eng.buffer("da1").float().cpu().t() @ xe + eng.buffer("Qb").float().cpu() @ W1.bfloat16().float()
    print("dW1 vs own-intermediates closed form", rel(t.disc.discriminator[0][0].weight.grad, dW1_chk))
    gW1 = t.disc.discriminator[0][0].weight.grad
    print("dW1 total vs closed", rel(gW1, dW1_loss + dW1_gp))
    print("dW1 total vs oracle", rel(gW1, o.disc.discriminator[0][0].weight.grad))
    print("closed vs oracle", rel(dW1_loss + dW1_gp, o.disc.discriminator[0][0].weight.grad))
    print("dW1 - loss part vs gp part", rel(gW1.cpu() - dW1_loss, dW1_gp), rel(gW1.cpu() - dW1_gp, dW1_loss))
    print("|loss part| |gp part|", dW1_loss.abs().max().item(), dW1_gp.abs().max().item())
    gW2 = t.disc.discriminator[1][0].weight.grad
    print("dW2 vs oracle", rel(gW2, o.disc.discriminator[1][0].weight.grad))
    print("dw3 vs oracle", rel(t.disc.final_layer.weight.grad, o.disc.final_layer.weight.grad))
    print("db1 vs oracle", rel(t.disc.discriminator[0][0].bias.grad, o.disc.discriminator[0][0].bias.grad))
    print("db2 vs oracle", rel(t.disc.discriminator[1][0].bias.grad, o.disc.discriminator[1][0].bias.grad))


if __name__ == "__main__":
    main()
