"""Per-tensor gradient agreement of one generator / critic step against the oracle (diagnostics; not a pytest file).

    python tests/gpu_debug_variant.py <variant> [small|mid] [gen|disc]
"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_parity as T  # noqa: E402
from oracle import restated  # noqa: E402

variant = sys.argv[1]
cfg = T.SMALL if (len(sys.argv) < 3 or sys.argv[2] == "small") else T.MID
which = sys.argv[3] if len(sys.argv) > 3 else "gen"
o, t = T.build_pair(variant, cfg, "adam")
B, G, L = cfg["B"], cfg["G"], cfg["latent"]
x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=6, ragged=True, text_dim=cfg["text_dim"],
                                   patch_dim=cfg["patch_dim"])
g = torch.Generator().manual_seed(3)
z = torch.randn(B, L, generator=g)
alpha = torch.rand(B, 1, generator=g)
dev = t.device
args = [c.to(dev) for c in T.ref_order(variant, x, cond)]
if which == "gen":
    o.train_gen(z, cond)
    t.train_gen(z.to(dev), *args)
    ref, got = o.gen, t.gen
else:
    o.train_disc(x, z, cond, alpha)
    t.train_disc(x.to(dev), z.to(dev), *args, alpha=alpha.to(dev))
    ref, got = o.disc, t.disc
torch.cuda.synchronize()
for (k, po), (_, pt) in zip(ref.named_parameters(), got.named_parameters()):
    if po.grad is None:
        print(f"{k:60s} ref None, got {'None' if pt.grad is None else 'tensor'}")
        continue
    d = (pt.grad.float().cpu() - po.grad).norm().item()
    n = po.grad.norm().item()
    print(f"{k:60s} |ref|={n:.3e} rel={d / max(n, 1e-30):.4f}")
# forward agreement: conditioning vectors, generator output, critic score
if variant != "vanilla":
    eng = t._engine(B)
    with torch.no_grad():
        o.gen.eval(); o.disc.eval()
        if variant in ("paper", "cross"):
            patches, ppad, text, tpad = cond
            cg = o.gen.conditioning(patches, ppad, text, tpad)
            cd = o.disc.conditioning(patches, ppad, text, tpad)
        else:
            text, patches, ppad = cond
            cg = o.gen.conditioning(patches, ppad, text)
            cd = o.disc.conditioning(patches, ppad, text)
    for name, refc in (("cond_gen", cg), ("cond_disc", cd)):
        gotc = eng.buffer(name)[:B].float().cpu()
        print(name, "max|ref|", refc.abs().max().item(), "max abs err", (gotc - refc).abs().max().item(),
              "rel fro", ((gotc - refc).norm() / refc.norm()).item())
