"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference/src via oracle/ref_shim.py) on CPU. Run in the build container:

    python -m oracle.make_golden

Each fixture holds the inputs, the explicit noise (z, alpha) and what the reference produced:
critic-step internals (generator output, critic scores, GP term, per-row gradient norms), every
parameter gradient after clipping, post-step weights, and loss curves over several train() calls.
Dropout is forced to 0 on the reference (its masks come from torch's RNG stream and are not
reproducible by any other implementation). Feature dims are reduced so the
fixtures stay small; one full-dims case per variant stores scalars only.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim, restated  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

SMALL = dict(B=8, G=203, P=5, T=3, embed=32, hidden=32, latent=16, text_dim=24, patch_dim=32)
FULL = dict(B=4, G=333, P=3, T=2, embed=256, hidden=256, latent=256, text_dim=768, patch_dim=1024)


def _digest(sd):
    import hashlib
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def ref_args(variant, x, cond):
    if variant in ("paper", "cross"):
        patches, ppad, text, tpad = cond
        return (x, text, tpad, patches, ppad)
    if variant in ("film", "concat", "concat_image", "img", "attn"):
        text, patches, ppad = cond
        return (x, text, patches, ppad)
    if variant == "label":
        return (x,) + tuple(cond)
    return (x,)


def model_args(variant, cond):
    """Argument order of the reference's generator/discriminator.forward after the first tensor."""
    return tuple(cond)


def draw_noise(seed, n_calls, n_critic, B, L):
    """Replays the RNG stream of reference train(): z, alpha, z, alpha, ..., z (per call)."""
    torch.manual_seed(seed)
    zs, alphas = [], []
    for _ in range(n_calls):
        for _ in range(n_critic):
            zs.append(torch.normal(0, 1, size=(B, L)))
            alphas.append(torch.rand(B, 1))
        zs.append(torch.normal(0, 1, size=(B, L)))
    return zs, alphas


def make(variant, cfg, optimizer, slope, n_calls, full_tensors, name):
    if variant == "attn":  # the reference's generator.forward prints its BatchNorm input / output on every call
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            _make(variant, cfg, optimizer, slope, n_calls, full_tensors, name)
        print(f"{name}: written")
        return
    _make(variant, cfg, optimizer, slope, n_calls, full_tensors, name)


def _make(variant, cfg, optimizer, slope, n_calls, full_tensors, name):
    c = dict(cfg)
    B, G, L = c["B"], c["G"], c["latent"]
    kw = dict(optimizer=optimizer, hidden=c["hidden"], latent=L, embed=c["embed"], seed=11,
              dropout=0.0, negative_slope=slope)
    if variant not in ("vanilla", "label"):
        kw.update(text_dim=c["text_dim"], patch_dim=c["patch_dim"])
    else:
        kw.pop("embed")   # (label: two 128-wide tables, hard-coded in the reference)
    t = ref_shim.make_trainer(variant, G, **kw) if variant != "vanilla" else \
        ref_shim.make_trainer(variant, G, optimizer=optimizer, hidden=c["hidden"], latent=L, seed=11,
                              dropout=0.0, negative_slope=slope)
    x, cond = restated.synthetic_batch(variant, B, G, c["P"], c["T"], seed=5, ragged=True,
                                       text_dim=c["text_dim"], patch_dim=c["patch_dim"])
    margs = model_args(variant, cond)
    zs, alphas = draw_noise(77, n_calls, t.n_critic, B, L)
    fx = dict(variant=variant, cfg=c, optimizer=optimizer, negative_slope=slope, init_seed=11,
              data_seed=5, noise_seed=77, n_calls=n_calls, x=x, cond=cond, zs=zs, alphas=alphas)
    # initial weights are reproducible from init_seed (same constructors, same order): keep digests
    fx["init_digest"] = dict(gen=_digest(t.gen.state_dict()), disc=_digest(t.disc.state_dict()))

    # --- first critic step, instrumented with the reference's own functions ---------------
    z0, a0 = zs[0], alphas[0]
    # (attn: the probes below are extra training-mode forwards the replayed train() calls do not contain; the
    # generator's BatchNorm running statistics are put back afterwards so that the fixture's final state is train()'s)
    buffers0 = {k: v.clone() for k, v in t.gen.named_buffers()}
    with torch.no_grad():
        fake = t.gen(z0, *margs)
        d_fake = t.disc(fake, *margs)
        d_true = t.disc(x, *margs)
    # gradient_penalty draws alpha = torch.rand(B,1) itself: replay the stream position
    torch.manual_seed(0)
    probe = torch.rand(B, 1)
    torch.manual_seed(0)
    for w in t.disc.parameters():
        w.requires_grad = True
    gp_probe = t.gradient_penalty(x, fake, *margs)
    fx["step0"] = dict(fake=fake, d_fake=d_fake, d_true=d_true, gp_alpha=probe, gp=gp_probe.detach())
    with torch.no_grad():
        for k, v in t.gen.named_buffers():
            v.copy_(buffers0[k])

    # --- the real thing: replay train() with the recorded noise ---------------------------
    torch.manual_seed(77)
    curves = dict(d=[], g=[])
    first = True
    for _ in range(n_calls):
        # reference train(): draws z/alpha from the global stream in the order draw_noise replays
        t.train(*ref_args(variant, x, cond)) if not first else None
        if first:
            # run the first call manually so the first critic / generator step can be captured
            xr = x.to(torch.float32)
            for i in range(t.n_critic):
                z = torch.normal(0, 1, size=(B, L))
                if variant == "vanilla":
                    t.train_disc(xr, z)
                elif variant in ("paper", "cross"):
                    t.train_disc(xr, z, *ref_args(variant, x, cond)[1:])
                else:
                    t.train_disc(xr, z, *ref_args(variant, x, cond)[1:])
                if i == 0:
                    fx["after_disc0"] = dict(
                        d_batch_loss=torch.tensor(t.d_batch_loss),
                        grads={k: (p.grad.clone() if p.grad is not None else None)
                               for k, p in t.disc.named_parameters()} if full_tensors else None,
                        grad_norms={k: (p.grad.norm().item() if p.grad is not None else None)
                                    for k, p in t.disc.named_parameters()})
            z = torch.normal(0, 1, size=(B, L))
            if variant == "vanilla":
                t.train_gen(z)
            else:
                t.train_gen(z, *ref_args(variant, x, cond)[1:])
            fx["after_gen0"] = dict(
                g_batch_loss=torch.tensor(t.g_batch_loss),
                grads={k: (p.grad.clone() if p.grad is not None else None)
                       for k, p in t.gen.named_parameters()} if full_tensors else None,
                grad_norms={k: (p.grad.norm().item() if p.grad is not None else None)
                            for k, p in t.gen.named_parameters()})
            first = False
        curves["d"].append(torch.tensor(t.d_batch_loss))
        curves["g"].append(torch.tensor(t.g_batch_loss))
    fx["curves"] = dict(d=torch.stack(curves["d"]), g=torch.stack(curves["g"]))
    fx["final_weight_norms"] = dict(
        gen={k: v.float().norm().item() for k, v in t.gen.state_dict().items()},
        disc={k: v.float().norm().item() for k, v in t.disc.state_dict().items()})
    if full_tensors:
        fx["final_gen"] = {k: v.clone() for k, v in t.gen.state_dict().items()}
        fx["final_disc"] = {k: v.clone() for k, v in t.disc.state_dict().items()}
    path = os.path.join(OUT, name + ".pt")
    torch.save(fx, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  d_loss0={fx['after_disc0']['d_batch_loss'].tolist()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]   # optional: variant names to (re)generate, e.g. `python -m oracle.make_golden label`
    if only:
        global make
        make_all = make
        make = lambda v, *a: make_all(v, *a) if v in only else None  # noqa: E731
    make("vanilla", SMALL, "adam", 0.0, 4, True, "vanilla_small_adam")
    make("vanilla", SMALL, "rms_prop", 0.2, 4, True, "vanilla_small_rmsprop_leaky")
    make("paper", SMALL, "adam", 0.0, 4, True, "paper_small_adam")
    make("paper", SMALL, "rms_prop", 0.0, 4, False, "paper_small_rmsprop")
    make("paper", SMALL, "adamw", 0.0, 3, False, "paper_small_adamw")
    make("film", SMALL, "adam", 0.0, 4, True, "film_small_adam")
    make("paper", FULL, "adam", 0.0, 3, False, "paper_fulldims_adam")
    make("film", FULL, "rms_prop", 0.0, 3, False, "film_fulldims_rmsprop")
    make("cross", SMALL, "adam", 0.0, 4, True, "cross_small_adam")
    make("img", SMALL, "rms_prop", 0.0, 4, True, "img_small_rmsprop")
    make("concat", SMALL, "adam", 0.0, 4, True, "concat_small_adam")
    make("concat_image", SMALL, "rms_prop", 0.2, 4, True, "concat_image_small_rmsprop_leaky")
    make("label", SMALL, "rms_prop", 0.0, 4, True, "label_small_rmsprop")
    make("label", SMALL, "adam", 0.2, 4, True, "label_small_adam_leaky")
    make("attn", SMALL, "adam", 0.0, 4, True, "attn_small_adam")
    make("attn", SMALL, "rms_prop", 0.2, 4, True, "attn_small_rmsprop_leaky")


if __name__ == "__main__":
    main()
