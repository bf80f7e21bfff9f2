"""GPU parity at the shapes BASELINE.json names and bench.py times (run with -m gpu on a B200), through the drop-in
classes and the C ABI, against the CPU oracle (oracle/restated.py, pinned to the unmodified reference):

  * cfg3 exactly (paper model, B=1024, G=18868, P=8, T=1): the split-K, CTA-pair and 256-wide-tile GEMM paths;
  * cfg2 tokens (film, P=256 -> 257-token self-attention inside the engine, G=18868);
  * cfg4 tokens (paper, P=64, T=32, G=20000, ragged masks): the mid-size attention + 32-key single-query attention;
  * post-step WEIGHTS after a full train() call (north_star: "post-step weights"), against the oracle and against
    the reference's own final weights stored in the golden fixtures;
  * the tcgen05 / TMA engine against the same engine on CUDA cores (GG_IMPL_SIMT_F32, same bf16 operands): what is
    left between the two is summation order, so the tensor-core path is exact to ~1e-3 and the allowance
    tests/test_gpu_parity.py makes against the fp32 oracle is bf16 rounding + ReLU mask flips, not a kernel bug;
  * a mask-free network (LeakyReLU slope 1.0 = identity): no unit can flip, and the gradients meet the bf16 bound
    directly;
  * dropout ON (the reference's and the bench's configuration): loss curves over 40 train() calls (240 optimizer
    steps) inside the oracle's own seed-to-seed spread.

Tolerances as stated in tests/test_gpu_parity.py: bf16 operands, fp32 accumulation against an
fp32 oracle -> 2e-2 of the tensor's scale on outputs / scores / GP; gradients by relative Frobenius error."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import restated
from test_gpu_parity import MID, TOL, build_pair, check_grads, fro, ref_order, rel

pytestmark = pytest.mark.gpu

FULLW = dict(embed=256, hidden=256, latent=256, text_dim=768, patch_dim=1024)
CFG3 = dict(B=1024, G=18868, P=8, T=1, **FULLW)                 # BASELINE.json configs[2], exactly
CFG2_TOKENS = dict(B=16, G=18868, P=256, T=1, **FULLW)          # configs[1] token count (script default --num_patches)
CFG4_TOKENS = dict(B=32, G=20000, P=64, T=32, **FULLW)          # configs[3] token counts


def _noise(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, L, generator=g), torch.rand(B, 1, generator=g), torch.randn(B, L, generator=g)


@pytest.mark.parametrize("variant,cfg,optimizer", [("paper", CFG3, "adam"), ("paper", CFG3, "rms_prop"),
                                                   ("film", CFG2_TOKENS, "adam"), ("paper", CFG4_TOKENS, "adam")])
def test_steps_at_baseline_shapes(variant, cfg, optimizer):
    """One train_disc + one train_gen (dropout 0) at a BASELINE.json shape: generator output, critic scores, per-row
    GP norms, GP, losses, every gradient (reference :376-461)."""
    o, t = build_pair(variant, cfg, optimizer)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=5, ragged=True)
    z, alpha, z2 = _noise(B, L, 99)
    dev = t.device
    args = [c.to(dev) for c in ref_order(variant, x, cond)]
    o.train_disc(x, z, cond, alpha)
    t.train_disc(x.to(dev), z.to(dev), *args, alpha=alpha.to(dev))
    eng = t._engine(B)
    torch.cuda.synchronize()
    assert rel(eng.buffer("fake_bf16"), o.last["fake"]) < TOL
    score = eng.buffer("score")[:, 0]
    assert rel(score[:B], o.last["d_fake"][:, 0]) < TOL
    assert rel(score[B:2 * B], o.last["d_true"][:, 0]) < TOL
    assert fro(eng.buffer("gp_norms")[:, 0], o.last["grad_norm"]) < TOL
    assert rel(eng.buffer("gp_norms")[:, 0], o.last["grad_norm"]) < 3 * TOL
    assert abs(t.last_gp - o.last["gp"].item()) <= TOL * max(abs(o.last["gp"].item()), 1e-3)
    np.testing.assert_allclose(t.d_batch_loss, o.d_batch_loss, rtol=TOL, atol=TOL * 0.05)
    check_grads([(k, p.grad) for k, p in o.disc.named_parameters()], list(t.disc.named_parameters()), "critic")
    # generator step on the critic each side has just updated itself (the updates differ by rounding only)
    o.train_gen(z2, cond)
    t.train_gen(z2.to(dev), *args)
    torch.cuda.synchronize()
    np.testing.assert_allclose(t.g_batch_loss, o.g_batch_loss, rtol=3 * TOL, atol=3 * TOL * 0.05)
    # (the generator step runs on the critic each side has just updated with its own rounding, and at 257 tokens
    # four ReLU / softmax layers sit upstream of most tower tensors: 0.102 measured for the film model at P=256)
    check_grads([(k, p.grad) for k, p in o.gen.named_parameters()], list(t.gen.named_parameters()), "generator",
                total=0.13)


def _update_agreement(before, ref_after, got_after):
    """Cosine and norm ratio of the two update vectors (w_after - w_before) over all trained tensors. Adam / RMSprop
    steps are lr * g / (|g| + eps)-like: an entry whose tiny gradient differs in sign moves by 2 lr, so updates are
    compared as vectors; the bound on single entries is the largest step the reference itself took."""
    dot = nr = ng = 0.0
    worst = 0.0
    for k in before:
        ur = (ref_after[k].float().cpu() - before[k]).flatten().double()
        ug = (got_after[k].float().cpu() - before[k]).flatten().double()
        if ur.abs().max().item() == 0.0:           # the never-used prototype layer (:114): untouched on both sides
            assert ug.abs().max().item() == 0.0, k
            continue
        dot, nr, ng = dot + (ur @ ug).item(), nr + (ur @ ur).item(), ng + (ug @ ug).item()
        worst = max(worst, (ug - ur).abs().max().item() / ur.abs().max().item())
    return dot / (nr * ng) ** 0.5, (ng / nr) ** 0.5, worst


@pytest.mark.parametrize("variant,optimizer", [("paper", "adam"), ("paper", "rms_prop"), ("film", "adam"),
                                               ("vanilla", "adam")])
def test_post_step_weights_after_one_train_call(variant, optimizer):
    """WGAN_GP.train (:463-477) = 5 critic steps + 1 generator step on recorded noise: the WEIGHTS of both nets after
    the call against the oracle's (clip + optimizer kernels on the engine's gradients, bf16 shadows refreshed between
    the steps)."""
    cfg = MID
    o, t = build_pair(variant, cfg, optimizer)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=7, ragged=True)
    g = torch.Generator().manual_seed(123)
    zs = [torch.randn(B, L, generator=g) for _ in range(6)]
    alphas = [torch.rand(B, 1, generator=g) for _ in range(5)]
    before_d = {k: v.clone() for k, v in o.disc.state_dict().items()}
    before_g = {k: v.clone() for k, v in o.gen.state_dict().items()}
    o.train(x, cond, zs, alphas)
    dev = t.device
    args = [c.to(dev) for c in ref_order(variant, x, cond)]
    t.train(x.to(dev), *args, zs=[z.to(dev) for z in zs], alphas=[a.to(dev) for a in alphas])
    torch.cuda.synchronize()
    stats = []
    for name, before, ref_sd, got_sd in (("critic", before_d, o.disc.state_dict(), t.disc.state_dict()),
                                         ("generator", before_g, o.gen.state_dict(), t.gen.state_dict())):
        cos, ratio, worst = _update_agreement(before, ref_sd, got_sd)
        print(f"{variant}/{optimizer} {name}: update cosine {cos:.4f}, norm ratio {ratio:.4f}, worst entry {worst:.3f}")
        stats.append((name, cos, ratio, worst))
    print(f"{variant}/{optimizer} losses got {t.d_batch_loss} {t.g_batch_loss} oracle {o.d_batch_loss} {o.g_batch_loss}")
    # RMSprop's first steps are 10 * lr * sign(g) for EVERY entry (v = 0.01 g^2), Adam's lr * sign(g): rounding in
    # near-zero gradient entries is amplified to full steps, so the losses after five such steps are compared loosely
    # for RMSprop (as tests/test_gpu_parity.py does for the golden loss curves) and the weights as update vectors
    rms = optimizer == "rms_prop"
    scale = max(np.abs(o.d_batch_loss).max(), abs(o.g_batch_loss[0]), 0.25)
    tol = (0.35 if rms else 0.05) * scale + 5e-3
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= tol and abs(t.g_batch_loss[0] - o.g_batch_loss[0]) <= tol
    # measured on the B200 (MID shapes): critic 0.98 / 0.97 / 0.995 (paper / film / vanilla, Adam), 0.86 (paper, RMSprop);
    # generator 0.87 - 0.98: its ONE step is lr * sign(g) of a gradient that came through a critic both sides had
    # already updated five times with their own rounding, i.e. the cosine is the share of entries whose sign agrees
    for name, cos, ratio, worst in stats:
        floor = 0.80 if (rms or name == "generator") else 0.93
        assert cos > floor and 0.9 < ratio < 1.1, (name, cos, ratio)
        assert worst <= 2.05, (name, worst)        # no entry moved further than a full sign flip of the largest step


@pytest.mark.parametrize("name", [n for n in golden_names() if "small" in n])
def test_final_weights_against_reference_golden(name):
    """The weights the UNMODIFIED reference ended with after n_calls train() calls (tests/golden, final_gen /
    final_disc / final_weight_norms) against the CUDA path replaying the recorded noise."""
    fx = load_golden(name)
    cfg, variant = fx["cfg"], fx["variant"]
    o, t = build_pair(variant, cfg, fx["optimizer"], fx["negative_slope"], seed=fx["init_seed"])
    dev = t.device
    x, cond, zs, alphas = fx["x"], fx["cond"], fx["zs"], fx["alphas"]
    args = [c.to(dev) for c in ref_order(variant, x, cond)]
    nc = t.n_critic
    for call in range(fx["n_calls"]):
        zc = [z.to(dev) for z in zs[call * (nc + 1):(call + 1) * (nc + 1)]]
        ac = [a.to(dev) for a in alphas[call * nc:(call + 1) * nc]]
        t.train(x.to(dev), *args, zs=zc, alphas=ac)
    torch.cuda.synchronize()
    for net, sd in (("gen", t.gen.state_dict()), ("disc", t.disc.state_dict())):
        for k, want in fx["final_weight_norms"][net].items():
            if variant == "attn" and net == "gen" and k in ("attention.in_proj_bias", "attention.out_proj.bias",
                                                            "patches_encoder.bias"):
                # constants in front of BatchNorm (and key biases under the softmax): the exact gradient is 0, what
                # RMSprop / Adam integrate — in the reference too — is round-off turned into +-lr-sized steps
                assert abs(sd[k].float().norm().item() - want) <= 0.1, (net, k)
                continue
            got = sd[k].float().norm().item()
            # (sign-like optimizer steps on tensors of 32 .. 8k entries: a flipped entry moves the norm by ~lr)
            assert abs(got - want) <= 1e-2 * want + 1.5e-2, (net, k, got, want)
    if fx.get("final_gen") is None:
        return
    for net, init, final_ref, sd in (("gen", o.gen.state_dict(), fx["final_gen"], t.gen.state_dict()),
                                     ("disc", o.disc.state_dict(), fx["final_disc"], t.disc.state_dict())):
        before = {k: v.clone() for k, v in init.items()}
        cos, ratio, worst = _update_agreement(before, final_ref, sd)
        print(f"{name} {net}: update cosine {cos:.4f}, norm ratio {ratio:.4f}, worst entry {worst:.3f}")
        # 8 x 32-unit nets, ~20 sign-like steps: measured on the B200 and pinned with margin
        assert cos > 0.80 and 0.85 < ratio < 1.15, (name, net, cos, ratio)


def _fresh_pair(variant, cfg, impl, seed=11):
    """Two drop-in trainers with identical weights, one per GEMM implementation."""
    _, t = build_pair(variant, cfg, "adam", seed=seed)
    t.gemm_impl = impl
    t._engines.clear()
    return t


@pytest.mark.parametrize("variant", ["paper", "vanilla"])
def test_tcgen05_engine_matches_cuda_core_engine(variant):
    """The same engine, same bf16 operands and activation storage, with every tcgen05 / TMA GEMM (and the grouped
    weight-gradient kernel) replaced by the CUDA-core fp32 check kernel (GG_IMPL_SIMT_F32): outputs and every gradient
    agree to summation order. This separates 'tensor-core path bug' from 'bf16 rounding / ReLU mask flip', which is all
    that remains in the comparisons with the fp32 oracle (the CUDA-core engine itself is checked against the oracle
    on the CPU suite, tests/test_engine_emulated.py)."""
    from gemmgan_b200 import _lib

    cfg = MID
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=5, ragged=True)
    z, alpha, z2 = _noise(B, L, 99)
    out = {}
    for impl in (_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_F32):
        t = _fresh_pair(variant, cfg, impl)
        dev = t.device
        args = [c.to(dev) for c in ref_order(variant, x, cond)]
        if variant == "vanilla":
            eng = t._engine(B)
            eng.set_batch(genes=x.to(dev))
        else:
            eng = t._stage(x.to(dev), *args)
        eng.disc_grads(z.to(dev), alpha.to(dev), training=False)
        torch.cuda.synchronize()
        rec = dict(fake=eng.buffer("fake_bf16").float().clone(), score=eng.buffer("score")[:2 * B, 0].clone(),
                   norms=eng.buffer("gp_norms")[:, 0].clone(), dgrads=t._flat_disc.grads.clone())
        eng.gen_grads(z2.to(dev), training=False)
        torch.cuda.synchronize()
        rec["ggrads"] = t._flat_gen.grads.clone()
        rec["slots"] = (dict(t._flat_disc.offsets), {s: p.numel() for s, p in t._flat_disc.slots.items()},
                        dict(t._flat_gen.offsets), {s: p.numel() for s, p in t._flat_gen.slots.items()})
        out[impl] = rec
    a, b = out[_lib.IMPL_TCGEN05], out[_lib.IMPL_SIMT_F32]
    m = dict(fake=rel(a["fake"], b["fake"]), score=rel(a["score"], b["score"]), norms=rel(a["norms"], b["norms"]))
    doff, dnum, goff, gnum = a["slots"]
    worst = {}
    for key, off, num in (("dgrads", doff, dnum), ("ggrads", goff, gnum)):
        w = 0.0
        for slot, o0 in off.items():
            ga, gb = a[key][o0:o0 + num[slot]], b[key][o0:o0 + num[slot]]
            n = gb.norm().item()
            if n > 1e-7 and num[slot] >= 4096:
                w = max(w, (ga - gb).norm().item() / n)
        worst[key] = w
        m[key] = fro(a[key], b[key])
    print(f"{variant}: tcgen05 vs CUDA-core engine {m}, worst tensor {worst}")
    # Activations are STORED in bf16 between kernels on both paths, so a different summation order moves some of them
    # by one bf16 ulp (2^-8 relative) and a unit within that distance of 0 may change ReLU branch: the two engines
    # agree to a few bf16 ulps on outputs, and several times tighter than either does with the fp32 oracle on gradients
    # measured: vanilla (trunk GEMMs, Gram-matrix GP, grouped weight gradients only) 3e-5 .. 3e-4 everywhere -- the
    # tcgen05 / TMA path is exact; paper (softmax, LayerNorm and three ReLU layers over bf16 activations upstream of most
    # tensors) 1.2 % / 2.6 % over all critic / generator gradients, 5 - 7 % on the worst tensor
    if variant == "vanilla":
        assert max(m.values()) < 1e-3 and max(worst.values()) < 1e-3, (m, worst)
    else:
        assert m["fake"] < 1e-2 and m["score"] < 1e-2 and m["norms"] < 6e-2, m   # (norms: worst single row)
        assert m["dgrads"] < 4e-2 and m["ggrads"] < 5e-2, m
        assert worst["dgrads"] < 0.10 and worst["ggrads"] < 0.12, worst


def test_gradients_meet_the_bf16_bound_without_relu_masks():
    """LeakyReLU(negative_slope=1.0) is the identity: the unconditional critic / generator become linear maps, no
    unit can change branch, and every gradient tensor is within 3 % (relative Frobenius; 0.2 - 2 % measured) of the fp32
    oracle — against the 15 % allowance tests/test_gpu_parity.py needs per tensor when ReLU masks can flip."""
    cfg = dict(MID, B=256, G=5000)
    o, t = build_pair("vanilla", cfg, "adam", slope=1.0)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch("vanilla", B, G, seed=5)
    z, alpha, z2 = _noise(B, L, 99)
    dev = t.device
    o.train_disc(x, z, cond, alpha)
    t.train_disc(x.to(dev), z.to(dev), alpha=alpha.to(dev))
    torch.cuda.synchronize()
    for (k, po), (_, pt) in zip(o.disc.named_parameters(), t.disc.named_parameters()):
        if po.grad.norm().item() > 1e-7:
            print("critic", k, fro(pt.grad, po.grad))
            assert fro(pt.grad, po.grad) < 3e-2, (k, fro(pt.grad, po.grad))
    o.train_gen(z2, cond)
    t.train_gen(z2.to(dev))
    torch.cuda.synchronize()
    for (k, po), (_, pt) in zip(o.gen.named_parameters(), t.gen.named_parameters()):
        print("generator", k, fro(pt.grad, po.grad))
        # (two bf16 roundings upstream of the generator: the critic's bf16 dD/dfake [B, G] and the bf16 activations)
        assert fro(pt.grad, po.grad) < 3e-2, (k, fro(pt.grad, po.grad))


def test_dropout_on_loss_curves_inside_the_oracle_spread():
    """Dropout 0.1 (live in train mode in the reference, :114-116, and in bench.py's configuration) draws from a
    different random stream than torch's: parity is statistical. 40 train() calls (240 optimizer steps) of the paper
    model on fixed data and fixed z / alpha noise; the oracle is run under 6 dropout seeds, the CUDA path under its own
    Philox stream. The CUDA curves must lie inside mean +- 4 sigma of the oracle's seed-to-seed spread (plus the bf16
    tolerance) at >= 90 % of the calls, and their averages over the run must agree."""
    cfg = dict(MID, B=32, G=500)
    variant, n_calls, n_seeds = "paper", 40, 6
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=7, ragged=True)
    g = torch.Generator().manual_seed(321)
    zs = [[torch.randn(B, L, generator=g) for _ in range(6)] for _ in range(n_calls)]
    alphas = [[torch.rand(B, 1, generator=g) for _ in range(5)] for _ in range(n_calls)]
    curves = []
    for s in range(n_seeds):
        torch.manual_seed(11)
        o = restated.OracleWGANGP(variant, G, optimizer="adam", dropout=None)   # dropout 0.1 as shipped
        torch.manual_seed(1000 + s)                                             # the dropout stream
        c = []
        for i in range(n_calls):
            o.train(x, cond, zs[i], alphas[i])
            c.append([o.d_batch_loss[0], o.d_batch_loss[1], o.d_batch_loss[2], o.g_batch_loss[0]])
        curves.append(c)
    ref = np.asarray(curves)                                                    # [seed, call, 4]
    mean, sd = ref.mean(0), ref.std(0, ddof=1)
    _, t = build_pair(variant, cfg, "adam", dropout=0.1)
    dev = t.device
    args = [c.to(dev) for c in ref_order(variant, x, cond)]
    got = []
    for i in range(n_calls):
        t.train(x.to(dev), *args, zs=[z.to(dev) for z in zs[i]], alphas=[a.to(dev) for a in alphas[i]])
        got.append([*t.d_batch_loss, t.g_batch_loss[0]])
    got = np.asarray(got)
    scale = np.abs(mean).max(0)
    band = 4.0 * sd + TOL * scale + 5e-3
    inside = np.abs(got - mean) <= band
    print("fraction of calls inside the band (d, d_real, d_fake, g):", inside.mean(0))
    print("run averages got / oracle:", got.mean(0), mean.mean(0), "mean sigma:", sd.mean(0))
    assert (inside.mean(0) >= 0.90).all(), inside.mean(0)
    assert np.all(np.abs(got.mean(0) - mean.mean(0)) <= 4.0 * sd.mean(0) / np.sqrt(n_seeds) + TOL * scale + 5e-3)
    # and dropout really is on: the dropout-free oracle sits measurably elsewhere or the spread is non-zero
    assert sd.mean() > 0


@pytest.mark.parametrize("variant,dropout", [("paper", 0.1), ("paper", 0.0), ("film", 0.1)])
def test_fused_encoder_layer_engine_matches_the_seven_launch_path(variant, dropout, monkeypatch):
    """The engine with the fused encoder-layer kernel (enc_layer.cu, default) against the same engine running every
    layer as seven launches (GEMMGAN_FUSED_LAYER=0), DROPOUT ON: both draw the same Philox masks (same seed, step,
    site and element indices), so the critic step must agree to bf16 rounding — conditioning vectors of all three
    replicas, scores, GP, losses and every gradient."""
    cfg = dict(MID, B=48)
    B, G, L = cfg["B"], cfg["G"], cfg["latent"]
    x, cond = restated.synthetic_batch(variant, B, G, cfg["P"], cfg["T"], seed=5, ragged=True)
    z, alpha, z2 = _noise(B, L, 99)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("GEMMGAN_FUSED_LAYER", fused)
        _, t = build_pair(variant, cfg, "adam", dropout=dropout)
        dev = t.device
        args = [c.to(dev) for c in ref_order(variant, x, cond)]
        eng = t._stage(x.to(dev), *args)
        eng.disc_grads(z.to(dev), alpha.to(dev), training=True)
        torch.cuda.synchronize()
        rec = dict(cond=eng.buffer("cond_disc").float().clone(), score=eng.buffer("score")[:, 0].clone(),
                   norms=eng.buffer("gp_norms")[:, 0].clone(), stats=eng.stats.clone(), dg=t._flat_disc.grads.clone())
        eng.gen_grads(z2.to(dev), training=True)
        torch.cuda.synchronize()
        rec["gg"] = t._flat_gen.grads.clone()
        rec["gstats"] = eng.stats.clone()
        out[fused] = rec
    a, b = out["1"], out["0"]
    m = {k: rel(a[k], b[k]) for k in ("cond", "score", "norms")}
    m["dgrads"], m["ggrads"] = fro(a["dg"], b["dg"]), fro(a["gg"], b["gg"])
    print(f"{variant} p={dropout}: fused vs seven-launch engine {m}")
    assert m["cond"] < 2e-2 and m["score"] < 2e-2 and m["norms"] < 6e-2, m
    assert m["dgrads"] < 5e-2 and m["ggrads"] < 8e-2, m   # (0.014 - 0.028 / 0.036 - 0.058 measured)
    assert torch.allclose(a["stats"][:5], b["stats"][:5], rtol=2e-2, atol=2e-3)
    assert torch.allclose(a["gstats"][4], b["gstats"][4], rtol=3e-2, atol=3e-3)


@pytest.mark.parametrize("B,G", [(256, 1000), (512, 5000)])
def test_gp_step_value_and_gradients_match_autograd(B, G):
    """gg_engine_gp_step (BASELINE config 5: the gradient penalty alone, value + gp_weight * dGP/d{W1, W2, w3}) on the
    unconditional critic against float64 autograd (gradient_penalty + the GP part of disc_loss.backward(),
    src/vanilla_gan_unconditional.py:304-327, :381). Critic layer 1 reads the fp32 real / fake profiles in place on the
    tensor cores as TF32; everything downstream is the Gram-matrix formulation on bf16 operands."""
    import vanilla_gan_unconditional as m

    torch.manual_seed(3)
    t = m.WGAN_GP_nocond(input_dims=G, latent_dims=256, vocab_sizes=[], generator_dims=[256, 256, G],
                         discriminator_dims=[256, 256, 1], optimizer="adam")
    t.build_WGAN_GP_nocond()
    t.init_train()
    dev = t.device
    g = torch.Generator(device=dev).manual_seed(1)
    real = torch.randn(B, G, device=dev, generator=g)
    fake = torch.randn(B, G, device=dev, generator=g)
    alpha = torch.rand(B, 1, device=dev, generator=g)
    eng = t._engine(B)
    gp = torch.zeros((), device=dev)
    t._flat_disc.grads.zero_()
    eng.gp_step(real, fake, alpha, gp)
    torch.cuda.synchronize()
    d = t.disc
    W1 = d.discriminator[0][0].weight.detach().double().requires_grad_(True)
    b1 = d.discriminator[0][0].bias.detach().double()
    W2 = d.discriminator[1][0].weight.detach().double().requires_grad_(True)
    b2 = d.discriminator[1][0].bias.detach().double()
    w3 = d.final_layer.weight.detach().double().requires_grad_(True)
    a = alpha.double()
    xh = (a * real.double() + (1 - a) * fake.double()).requires_grad_(True)
    out = torch.relu(torch.relu(xh @ W1.t() + b1) @ W2.t() + b2) @ w3.t()
    (gr,) = torch.autograd.grad(out.sum(), xh, create_graph=True)
    ref = ((gr.norm(dim=1) - 1) ** 2).mean()
    (10.0 * ref).backward()
    assert abs(gp.item() - ref.item()) <= 1e-2 * abs(ref.item())
    for name, p_ref, p_got in (("W1", W1, d.discriminator[0][0].weight), ("W2", W2, d.discriminator[1][0].weight),
                               ("w3", w3, d.final_layer.weight)):
        e = fro(p_got.grad, p_ref.grad.float())
        print(f"gp_step B={B} G={G} {name}: rel fro {e:.4f}")
        assert e < 6e-2, (name, e)          # ReLU mask flips at bf16 / TF32 precision, as in check_grads
