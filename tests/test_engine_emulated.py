"""The WHOLE training step on the CPU suite: gg_engine_* (gemmgan_b200/csrc/engine.cu and every kernel file it
sequences) compiled for the host (tests/cuda_emu/emu_engine.cpp) in the engine's all-CUDA-core configuration
(gemm_impl = GG_IMPL_SIMT_F32), driven through gemmgan_b200/runtime.py exactly as the trainer drives it
(set_batch -> disc_grads -> optim_step, gen_grads -> optim_step), against the oracle (oracle/restated.py, pinned to the
unmodified reference) on the same weights, batch and noise: WGAN_GP.train_disc / train_gen of
src/conditional_gan_cross_attention_with_film.py:376-461 and the same steps of every other variant (vanilla, film,
cross, img, concat text / image, label).

What this covers that the per-kernel emulation tests do not: the hand-written backward / double backward as the engine
sequences it (Gram-matrix gradient penalty, shared [fake; real] GEMM, replica batching, FiLM and tower backward, two
post-norm encoder layers, both single-query attentions), the parameter slot tables, the clip + optimizer update on the
flat buffers. What it cannot cover: the tcgen05 / TMA GEMMs and the grouped weight-gradient kernel (GPU only; on the
B200 tests/test_gpu_parity.py makes the same comparisons through the drop-in trainers), lanes, graphs, NCCL.

Tolerances are those of tests/test_gpu_parity.py (bf16 operands, fp32 accumulation): 2e-2 of the tensor's scale on
outputs / scores / GP, relative Frobenius on gradients (ReLU mask flips).
"""
import contextlib
import ctypes as C

import numpy as np
import pytest
import torch

import emu_build
from gemmgan_b200 import _abi_decl as A
from gemmgan_b200 import _lib, runtime
from oracle import restated

TOL = 2e-2
SMALL = dict(B=8, G=203, P=5, T=3, embed=32, hidden=32, latent=16, text_dim=24, patch_dim=32)
# the reference's feature widths (768-d text, 1024-d patches, E = H = 256 -> 4 heads of 64): the production kernel paths
# (register-resident / mma.sync attention, vectorised LayerNorm and FiLM) instead of the generic small-width ones
FULL = dict(B=16, G=1000, P=8, T=2, embed=256, hidden=256, latent=256, text_dim=768, patch_dim=1024)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("engine", tmp_path_factory.mktemp("cuda_emu"), cudart=True)
    A.declare(L)
    return L


@pytest.fixture()
def rt(emu, monkeypatch):
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setattr(_lib, "require_device", lambda dev=0: None)
    monkeypatch.setattr(runtime, "_stream", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    return runtime


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def fro(a, b):
    a, b = a.detach().float(), b.detach().float()
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


def build(rt, variant, cfg, optimizer, slope, seed=11):
    """(oracle, generator, critic, engine) with identical initial weights; nets from the drop-in modules."""
    torch.manual_seed(seed)
    o = restated.OracleWGANGP(variant, cfg["G"], latent=cfg["latent"], embed=cfg["embed"], hidden=cfg["hidden"],
                              optimizer=optimizer, negative_slope=slope, dropout=0.0, text_dim=cfg["text_dim"],
                              patch_dim=cfg["patch_dim"])
    torch.manual_seed(seed)
    H, G = cfg["hidden"], cfg["G"]
    if variant == "vanilla":
        import vanilla_gan_unconditional as m
        gen, disc = m.WGAN_GP_model_nocond(cfg["latent"], G, [], [], [H, H, G], [H, H, 1], slope, False)
        shape = dict(E=0, H=H, Dt=0, Dp=0, P=0, T=0)
        clip_d = clip_g = 0.0
    elif variant in ("concat", "concat_image"):
        import conditional_gan_concat as m
        image = variant == "concat_image"
        din = cfg["patch_dim"] if image else cfg["text_dim"]
        gen, disc = m.WGAN_GP_model(cfg["latent"], G, din, cfg["embed"], [H, H, G], [H, H, 1],
                                    "image" if image else "text", slope, False)
        shape = dict(E=cfg["embed"], H=H, Dt=din, Dp=din, P=1, T=1, tower_bias=True)
        clip_d, clip_g = float(m.WGAN_GP.clip_d or 0.0), float(m.WGAN_GP.clip_g or 0.0)
    elif variant == "label":
        import benchmark_generative_model as m
        gen, disc = m.WGAN_GP_model_benchmark(cfg["latent"], G, [], [10, 10], [H, H, G], [H, H, 1], slope, False)
        shape = dict(E=gen.categorical_embedded_dims, H=H, Dt=10, Dp=10, P=1, T=1)
        clip_d, clip_g = float(m.WGAN_GP_benchmark.clip_d or 0.0), float(m.WGAN_GP_benchmark.clip_g or 0.0)
    else:
        import importlib
        m = importlib.import_module({"paper": "conditional_gan_cross_attention_with_film", "film": "conditional_gan_film",
                                     "cross": "conditional_gan_cross_attention",
                                     "img": "conditional_gan_img_transformer",
                                     "attn": "conditional_gan_attention"}[variant])
        gen, disc = m.WGAN_GP_model(cfg["latent"], G, cfg["embed"], [H, H, G], [H, H, 1], cfg["text_dim"],
                                    cfg["patch_dim"], slope, False)
        tokens = variant in ("paper", "cross")           # text tokens [B, T, Dt] vs one text embedding [B, Dt]
        shape = dict(E=cfg["embed"], H=H, Dt=cfg["text_dim"], Dp=cfg["patch_dim"], P=cfg["P"],
                     T=cfg["T"] if tokens else 1, tower_bias=variant in ("paper", "attn"))
        clip_d, clip_g = float(m.WGAN_GP.clip_d or 0.0), float(m.WGAN_GP.clip_g or 0.0)
    for (k1, v1), (k2, v2) in zip(o.gen.state_dict().items(), gen.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    for (k1, v1), (k2, v2) in zip(o.disc.state_dict().items(), disc.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    dev = torch.device("cpu")
    fg, fd = rt.FlatNet(gen, dev, optimizer), rt.FlatNet(disc, dev, optimizer)
    eng = rt.Engine(variant="concat" if variant.startswith("concat") else variant, B=cfg["B"], G=G, L=cfg["latent"], gen=fg, disc=fd, slope=slope, dropout_p=0.0,
                    gp_weight=10.0, clip_d=clip_d, clip_g=clip_g, optimizer=optimizer, gemm_impl=_lib.IMPL_SIMT_F32,
                    device=dev, **shape)
    eng.set_lanes(False)
    return o, gen, disc, eng


def stage(eng, variant, x, cond):
    if variant == "vanilla":
        eng.set_batch(genes=x)
    elif variant == "label":
        eng.set_batch(genes=x)
        eng.set_labels(cond[0], cond[1])
    elif variant == "concat":
        eng.set_batch(genes=x, text=cond[0])
    elif variant == "concat_image":     # masked mean of the patch embeddings first, then ONE encoder GEMM
        eng.set_batch(genes=x, text=eng.masked_mean_rows(cond[1], cond[2]))
    elif variant in ("paper", "cross"):
        patches, ppad, text, tpad = cond
        eng.set_batch(genes=x, patches=patches, patch_pad=ppad, text=text, text_pad=tpad)
    else:
        text, patches, ppad = cond
        eng.set_batch(genes=x, patches=patches, patch_pad=ppad, text=text, text_pad=None)


def check_grads(named_ref, named_got, total):
    num = den = 0.0
    for (k, gref), (_, pt) in zip(named_ref, named_got):
        if gref is None:
            assert pt.grad is None, k
            continue
        got = pt.grad.detach().float()
        assert torch.isfinite(got).all(), k
        d, n = (got - gref).norm().item(), gref.norm().item()
        num, den = num + d * d, den + n * n
        if n > 1e-7:
            assert d / n <= (0.15 if gref.numel() >= 4096 else 0.35), (k, d / n)
    assert (num / max(den, 1e-30)) ** 0.5 <= total


@pytest.mark.parametrize("variant,optimizer,slope", [("vanilla", "adam", 0.0), ("vanilla", "rms_prop", 0.2),
                                                     ("paper", "adam", 0.0), ("paper", "rms_prop", 0.0),
                                                     ("film", "adam", 0.0), ("cross", "adam", 0.0), ("img", "adam", 0.0),
                                                     ("concat", "adam", 0.2), ("concat_image", "rms_prop", 0.0),
                                                     ("label", "adam", 0.0), ("attn", "adam", 0.0),
                                                     ("attn", "rms_prop", 0.2)])
def test_critic_and_generator_step_match_the_oracle(rt, variant, optimizer, slope):
    run_steps(rt, variant, optimizer, slope, SMALL)


@pytest.mark.parametrize("variant,P", [("paper", 8), ("film", 20)])
def test_steps_at_the_reference_feature_widths(rt, variant, P):
    """paper: 8 patches + CLS = 9 tokens (short attention kernel, cfg3's shape); film: 20 patches + CLS = 21 tokens
    (mma.sync mid kernel)."""
    run_steps(rt, variant, "adam", 0.0, dict(FULL, P=P))


@pytest.mark.parametrize("variant", ["paper", "cross"])
@pytest.mark.parametrize("shortcut", ["1", "0"])
def test_single_text_token_step(rt, monkeypatch, variant, shortcut):
    """T = 1 (BASELINE configs 1-3): the text2patch attention is a softmax over one key. With GEMMGAN_T1_SHORTCUT (default)
    the engine forms c = pv + (v_text Wo^T + bo) directly and sends no gradient to the query / key projections (exact
    zeros, as autograd produces); the general path ("0") runs the attention kernels. Both against the oracle."""
    monkeypatch.setenv("GEMMGAN_T1_SHORTCUT", shortcut)
    run_steps(rt, variant, "adam", 0.0, dict(SMALL, T=1))


def test_cfg4_like_token_counts(rt):
    """64 patches + CLS = 65 tokens (mma.sync mid kernel) and 32 text tokens (single-query attention over 65 / 32 keys):
    BASELINE config 4's sequence lengths at a batch the emulation finishes quickly."""
    run_steps(rt, "paper", "adam", 0.0, dict(FULL, B=8, P=64, T=32))


def run_steps(rt, variant, optimizer, slope, cfg):
    o, gen, disc, eng = build(rt, variant, cfg, optimizer, slope)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=5, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    g = torch.Generator().manual_seed(99)
    z, alpha, z2 = torch.randn(B, L, generator=g), torch.rand(B, 1, generator=g), torch.randn(B, L, generator=g)
    lr = 5e-4

    # ---- critic step (train_disc, :376-423)
    o.train_disc(x, z, cond, alpha)
    stage(eng, variant, x, cond)
    eng.disc_grads(z, alpha, training=True)
    assert rel(eng.buffer("fake_bf16"), o.last["fake"]) < TOL
    score = eng.buffer("score")[:, 0]
    assert rel(score[:B], o.last["d_fake"][:, 0]) < TOL
    assert rel(score[B:2 * B], o.last["d_true"][:, 0]) < TOL
    assert fro(eng.buffer("gp_norms")[:, 0], o.last["grad_norm"]) < TOL          # per-row ||dD/dx_hat||
    st = eng.stats
    assert abs(st[A.STAT_GP].item() - o.last["gp"].item()) <= TOL * max(abs(o.last["gp"].item()), 1e-3)
    got_d = np.array([st[A.STAT_LOSS_REAL].item() + st[A.STAT_LOSS_FAKE].item(), st[A.STAT_LOSS_REAL].item(),
                      st[A.STAT_LOSS_FAKE].item()])
    np.testing.assert_allclose(got_d, o.d_batch_loss, rtol=TOL, atol=TOL * 0.05)
    before = {k: p.detach().clone() for k, p in disc.named_parameters()}
    eng.optim_step(A.NET_DISC, lr)                                               # clip (:414) + optimizer (:415)
    check_grads([(k, p.grad) for k, p in o.disc.named_parameters()], list(disc.named_parameters()), total=0.08)
    # post-step weights: the update vectors of both sides (Adam / RMSprop steps are sign-like: an element whose tiny
    # gradient differs in sign moves by 2 lr, so the updates are compared as vectors, not element by element)
    dot = nr = ng = 0.0
    for (k, po), (_, pt) in zip(o.disc.named_parameters(), disc.named_parameters()):
        if po.grad is None:                                                      # the unused prototype layer (:114)
            assert torch.equal(pt.detach(), before[k]), k
            continue
        ur, ug = (po.detach() - before[k]).flatten(), (pt.detach() - before[k]).flatten()
        assert ug.abs().max().item() <= 1.01 * max(ur.abs().max().item(), lr), k
        dot, nr, ng = dot + (ur @ ug).item(), nr + (ur @ ur).item(), ng + (ug @ ug).item()
    # (Adam's first step is lr * sign(g): every near-zero gradient entry whose sign differs costs 2 lr; 0.956 seen for img)
    assert dot / (nr * ng) ** 0.5 > 0.9 and 0.9 < (ng / nr) ** 0.5 < 1.1
    for (k, po), (_, pt) in zip(o.gen.named_parameters(), gen.named_parameters()):
        assert torch.equal(po.detach(), pt.detach()), k                          # generator untouched

    # ---- generator step (train_gen, :425-461), on the critic the kernels have just updated
    with torch.no_grad():                                                        # same critic on both sides
        for po, pt in zip(o.disc.parameters(), disc.parameters()):
            po.copy_(pt)
    o.train_gen(z2, cond)
    eng.sync_external_param_writes()
    eng.gen_grads(z2, training=True)
    assert st[A.STAT_G_LOSS].item() == pytest.approx(float(np.asarray(o.g_batch_loss).reshape(-1)[0]), rel=TOL, abs=TOL * 0.05)
    eng.optim_step(A.NET_GEN, lr)
    # img at this size: one ReLU flip among the 8 x 32 units of the patch encoder moves every gradient by ~1/8
    # (tests/test_gpu_parity.py makes the same exception and runs its generator step at the larger configuration)
    check_grads([(k, p.grad) for k, p in o.gen.named_parameters()], list(gen.named_parameters()),
                total=0.2 if variant == "img" else 0.08)
    if variant == "attn":   # BatchNorm1d running statistics after the two training-mode generator forwards (:108, :126)
        bo, bt = o.gen.attn_bn, gen.attn_bn
        assert not torch.equal(bt.running_mean, torch.zeros_like(bt.running_mean))
        torch.testing.assert_close(bt.running_mean, bo.running_mean, rtol=TOL, atol=2e-3)
        torch.testing.assert_close(bt.running_var, bo.running_var, rtol=TOL, atol=2e-3)


def test_generate_and_critic_entry_points(rt):
    """generate_samples (:601-608) and discriminator.forward as stand-alone calls (eval mode)."""
    cfg = SMALL
    o, gen, disc, eng = build(rt, "paper", cfg, "adam", 0.0)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch("paper", B, G, cfg["P"], cfg["T"], seed=6, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    z = torch.randn(B, L, generator=torch.Generator().manual_seed(3))
    stage(eng, "paper", x, cond)
    fake = eng.generate(z)
    with torch.no_grad():
        want = o.gen(z, *cond)
        want_score = o.disc(x, *cond)
    assert rel(fake, want) < TOL
    assert rel(eng.critic(x), want_score) < TOL
    # WGAN_GP.gradient_penalty (:351-374) as a stand-alone call
    alpha = torch.rand(B, 1, generator=torch.Generator().manual_seed(4))
    gp = eng.gradient_penalty(x, want, alpha, training=False)
    o.disc.eval()
    gp_ref = o.gradient_penalty(x, want, cond, alpha)
    assert gp.item() == pytest.approx(gp_ref.item(), rel=TOL)


def test_attention_variant_generates_from_the_running_statistics_in_eval_mode(rt):
    """generate_samples of conditional_gan_attention.py (:505-512, gen.eval()): BatchNorm1d normalises with the running
    statistics the training-mode forwards have accumulated and leaves them alone; a training-mode forward of the same
    call uses the batch statistics and updates them."""
    cfg = SMALL
    o, gen, disc, eng = build(rt, "attn", cfg, "adam", 0.0)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch("attn", B, G, cfg["P"], cfg["T"], seed=6, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    g = torch.Generator().manual_seed(3)
    stage(eng, "attn", x, cond)
    for _ in range(3):                                   # accumulate running statistics on both sides
        z = torch.randn(B, L, generator=g)
        with torch.no_grad():
            want = o.gen(z, *cond)
        assert rel(eng.generate(z, training=True), want) < TOL
    bo, bt = o.gen.attn_bn, gen.attn_bn
    torch.testing.assert_close(bt.running_mean, bo.running_mean, rtol=TOL, atol=2e-3)
    torch.testing.assert_close(bt.running_var, bo.running_var, rtol=TOL, atol=2e-3)
    frozen = (bt.running_mean.clone(), bt.running_var.clone())
    o.gen.eval()
    z = torch.randn(B, L, generator=g)
    with torch.no_grad():
        want = o.gen(z, *cond)
    assert rel(eng.generate(z, training=False), want) < TOL
    assert torch.equal(bt.running_mean, frozen[0]) and torch.equal(bt.running_var, frozen[1])
    assert rel(eng.critic(x), o.disc(x, *cond).detach()) < TOL


def test_dropout_is_deterministic_in_the_seed_and_close_in_expectation(rt):
    """Dropout on (the reference's p = 0.1): the masks come from the engine's Philox stream, not torch's, so the step is
    checked for reproducibility under a fixed (seed, step) and for staying near the dropout-free step."""
    cfg = SMALL
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch("paper", B, G, cfg["P"], cfg["T"], seed=5, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    g = torch.Generator().manual_seed(1)
    z, alpha = torch.randn(B, L, generator=g), torch.rand(B, 1, generator=g)

    def run(p):
        import conditional_gan_cross_attention_with_film as m
        torch.manual_seed(11)
        H = cfg["hidden"]
        gen, disc = m.WGAN_GP_model(L, G, cfg["embed"], [H, H, G], [H, H, 1], cfg["text_dim"], cfg["patch_dim"], 0.0, False)
        dev = torch.device("cpu")
        fg, fd = rt.FlatNet(gen, dev, "adam"), rt.FlatNet(disc, dev, "adam")
        eng = rt.Engine(variant="paper", B=B, G=G, L=L, gen=fg, disc=fd, slope=0.0, dropout_p=p, gp_weight=10.0,
                        clip_d=10.0, clip_g=2.0, optimizer="adam", gemm_impl=_lib.IMPL_SIMT_F32, device=dev, seed=7,
                        E=cfg["embed"], H=H, Dt=cfg["text_dim"], Dp=cfg["patch_dim"], P=cfg["P"], T=cfg["T"], tower_bias=True)
        eng.set_lanes(False)
        stage(eng, "paper", x, cond)
        eng.disc_grads(z, alpha, training=True)
        return eng.buffer("fake_bf16").float().clone(), fd.grads.clone(), eng.stats.clone()

    f1, g1, s1 = run(0.1)
    f2, g2, s2 = run(0.1)
    assert torch.equal(f1, f2) and torch.equal(g1, g2) and torch.equal(s1, s2)
    f0, g0, s0 = run(0.0)
    assert not torch.equal(f1, f0)
    assert rel(f1, f0) < 0.5 and fro(g1, g0) < 0.6           # perturbed, not different in kind


@pytest.mark.parametrize("name", ["vanilla_small_adam", "paper_small_adam", "film_small_adam", "label_small_rmsprop",
                                  "attn_small_adam"])
def test_against_reference_golden_loss_curves(rt, name):
    """Replays the recorded noise of a run of the UNMODIFIED reference (tests/golden, oracle/make_golden.py) through
    the emulated engine: first critic step against the recorded internals, then whole train() calls (n_critic critic
    steps + one generator step, :463-477) against the reference's loss curves."""
    from conftest import load_golden

    fx = load_golden(name)
    cfg, variant = fx["cfg"], fx["variant"]
    o, gen, disc, eng = build(rt, variant, cfg, fx["optimizer"], fx["negative_slope"], seed=fx["init_seed"])
    x, cond, zs, alphas = fx["x"], fx["cond"], fx["zs"], fx["alphas"]
    B, nc, lr = cfg["B"], o.n_critic, 5e-4
    st = eng.stats
    stage(eng, variant, x, cond)

    def critic_step(z, alpha):
        eng.disc_grads(z, alpha, training=True)
        d = np.array([st[A.STAT_LOSS_REAL].item() + st[A.STAT_LOSS_FAKE].item(), st[A.STAT_LOSS_REAL].item(),
                      st[A.STAT_LOSS_FAKE].item()])
        eng.optim_step(A.NET_DISC, lr)
        return d

    def gen_step(z):
        eng.gen_grads(z, training=True)
        g_loss = np.array([st[A.STAT_G_LOSS].item()])
        eng.optim_step(A.NET_GEN, lr)
        return g_loss

    eng.disc_grads(zs[0], alphas[0], training=True)
    assert rel(eng.buffer("fake_bf16"), fx["step0"]["fake"]) < TOL
    assert rel(eng.buffer("score")[:B, 0], fx["step0"]["d_fake"][:, 0]) < TOL
    assert rel(eng.buffer("score")[B:2 * B, 0], fx["step0"]["d_true"][:, 0]) < TOL
    d_curve, g_curve = [], []
    for call in range(fx["n_calls"]):
        d = None
        for i in range(nc):
            d = critic_step(zs[call * (nc + 1) + i], alphas[call * nc + i])
        d_curve.append(d)                              # d_batch_loss of the LAST critic step of the call (:421-423)
        g_curve.append(gen_step(zs[call * (nc + 1) + nc]))
    d_ref, g_ref = fx["curves"]["d"].numpy(), fx["curves"]["g"].numpy()
    rms = fx["optimizer"] == "rms_prop"
    for call in range(fx["n_calls"]):
        tol = (0.05 if call == 0 else 0.35) if rms else 0.05      # as in tests/test_gpu_parity.py
        scale = max(np.abs(d_ref[call]).max(), np.abs(g_ref[call]).max(), 0.25 if rms else 0.05)
        assert np.abs(d_curve[call] - d_ref[call]).max() <= tol * scale + 2e-3, (call, d_curve[call], d_ref[call])
        assert np.abs(g_curve[call] - g_ref[call]).max() <= tol * scale + 2e-3, (call, g_curve[call], g_ref[call])


def test_staged_backward_of_the_data_parallel_path_equals_the_whole_step(rt):
    """The data-parallel trainer cuts the backward into phases (trunk, cross-attention tail, encoder layers, embedding
    tail: gg_engine_*_grads_phase) so that each gradient bucket can be all-reduced while the next stage runs. The
    staged sequences must leave exactly the gradients of the one-call step."""
    cfg = SMALL
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch("paper", B, G, cfg["P"], cfg["T"], seed=5, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    g = torch.Generator().manual_seed(2)
    z, alpha = torch.randn(B, L, generator=g), torch.rand(B, 1, generator=g)

    def grads(schedule):
        _, gen, disc, eng = build(rt, "paper", cfg, "adam", 0.0)
        stage(eng, "paper", x, cond)
        for ph in schedule:
            eng.disc_grads(z, alpha, training=True, phase=ph)
        d = eng.disc.grads.clone()
        for ph in schedule:
            eng.gen_grads(z, training=True, phase=ph)
        return d, eng.gen.grads.clone()

    whole_d, whole_g = grads([0])
    two_d, two_g = grads([1, 2])                                             # trunk | tower
    nj = A.PHASE_NO_JOIN
    staged = [1 | nj] + [(A.PHASE_STAGE0 + s) | (0 if s == 3 else nj) for s in range(4)]
    st_d, st_g = grads(staged)                                               # trunk | head | layer | layer | embedding
    assert torch.equal(whole_d, two_d) and torch.equal(whole_g, two_g)
    assert torch.equal(whole_d, st_d) and torch.equal(whole_g, st_g)


def test_engines_of_other_batch_sizes_see_optimizer_updates(rt):
    """The reference's loaders have no drop_last (multi_patch_multi_token_gan_dataloader.py:178-185): the last partial
    batch of an epoch, and a validation loader with its own batch size, run on an engine of another B than the one
    whose optimizer kernel last moved the weights. Every engine keeps its own bf16 weight shadows: after a step on
    engine A, engine B must refresh (FlatNet.generation / Engine.sync_params) — compared with a fresh engine built
    from the updated weights, bit for bit."""
    cfg = SMALL
    o, gen, disc, eng = build(rt, "paper", cfg, "adam", 0.0)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    shape = dict(E=cfg["embed"], H=cfg["hidden"], Dt=cfg["text_dim"], Dp=cfg["patch_dim"], P=cfg["P"], T=cfg["T"],
                 tower_bias=True)

    def engine_at(b):
        e = rt.Engine(variant="paper", B=b, G=G, L=L, gen=eng.gen, disc=eng.disc, slope=0.0, dropout_p=0.0,
                      gp_weight=10.0, clip_d=10.0, clip_g=2.0, optimizer="adam", gemm_impl=_lib.IMPL_SIMT_F32,
                      device=torch.device("cpu"), **shape)
        e.set_lanes(False)
        return e

    small = engine_at(5)                                   # created BEFORE the updates: its shadows are the old weights
    x, cond = restated.synthetic_batch("paper", B, G, cfg["P"], cfg["T"], seed=5, ragged=True,
                                       text_dim=cfg["text_dim"], patch_dim=cfg["patch_dim"])
    g = torch.Generator().manual_seed(99)
    z, alpha, z2 = torch.randn(B, L, generator=g), torch.rand(B, 1, generator=g), torch.randn(B, L, generator=g)
    stage(eng, "paper", x, cond)
    eng.disc_grads(z, alpha, training=True)
    eng.optim_step(A.NET_DISC, 5e-2)                       # a large step: stale shadows would be far off
    eng.stepped(A.NET_DISC)
    eng.gen_grads(z2, training=True)
    eng.optim_step(A.NET_GEN, 5e-2)
    eng.stepped(A.NET_GEN)
    assert small.seen[A.NET_DISC] != eng.disc.generation and small.seen[A.NET_GEN] != eng.gen.generation
    xs, conds = x[:5], tuple(c[:5] for c in cond)
    zs = z[:5].contiguous()
    stage(small, "paper", xs, conds)
    stale_fake, stale_score = small.generate(zs).clone(), small.critic(xs).clone()
    small.sync_params()                                    # what TrainerBase._engine does for a cached engine
    assert small.seen[A.NET_DISC] == eng.disc.generation and small.seen[A.NET_GEN] == eng.gen.generation
    fresh = engine_at(5)
    stage(fresh, "paper", xs, conds)
    want_fake, want_score = fresh.generate(zs), fresh.critic(xs)
    got_fake, got_score = small.generate(zs), small.critic(xs)
    assert torch.equal(got_fake, want_fake) and torch.equal(got_score, want_score)
    assert not torch.equal(stale_fake, want_fake) and not torch.equal(stale_score, want_score)
    # a write through PyTorch (load_state_dict) is seen by every engine, not only the first one that asks
    with torch.no_grad():
        gen.final_layer.bias.add_(1.0)
        gen.final_layer.weight.mul_(0.5)
    for e in (eng, small, fresh):
        e.sync_params()
    stage(eng, "paper", x, cond)
    a = eng.generate(z)[:5]
    stage(small, "paper", xs, conds)
    b = small.generate(zs)
    assert torch.allclose(a, b, atol=0, rtol=0)
