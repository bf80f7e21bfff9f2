"""The drop-in trainer classes themselves on the CPU suite: `WGAN_GP` / `WGAN_GP_nocond` of the root modules
(gemmgan_b200/trainer.py::TrainerBase) driving the engine compiled for the host (tests/cuda_emu/emu_engine.cpp,
all-CUDA-core configuration), called the way a user of the reference calls them — build_WGAN_GP(), init_train(),
train(...), fit(loader), generate_samples(...), state_dict()s — against the oracle (oracle/restated.py, pinned to
the unmodified reference; src/conditional_gan_cross_attention_with_film.py:256-477,
src/vanilla_gan_unconditional.py:211-431) and against themselves.

What this adds to tests/test_engine_emulated.py (which drives runtime.Engine directly): the host logic between the
reference-facing methods and the engine — one engine per batch size sharing the flat buffers (a last partial batch, a
validation loader of another batch size), networks rebuilt by fit(), noise staged up front, LR halving through
`param_groups`, optimizer / network checkpoints. Not covered here (GPU only): CUDA graphs, lanes, prefetch, NCCL.
"""
import importlib
import io

import numpy as np
import pytest
import torch

import emu_build
import host_trainer
from gemmgan_b200 import _abi_decl as A
from oracle import ref_shim, restated

TOL = 2e-2
COS_FLOOR = 0.75       # see update_cosine()
SMALL = dict(G=203, P=5, T=3, embed=32, hidden=32, latent=16, text_dim=24, patch_dim=32)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("engine", tmp_path_factory.mktemp("cuda_emu"), cudart=True)
    A.declare(L)
    return L


@pytest.fixture()
def host(emu, monkeypatch):
    """The trainer on the host: the library is the emulated build, the device is the CPU, steps run eagerly."""
    return host_trainer.apply(monkeypatch.setattr, emu)


def make(variant, optimizer="adam", seed=11, **kw):
    """(oracle, drop-in trainer) holding the same initial weights."""
    c = SMALL
    H, G = c["hidden"], c["G"]
    torch.manual_seed(seed)
    o = restated.OracleWGANGP(variant, G, latent=c["latent"], embed=c["embed"], hidden=H, optimizer=optimizer,
                              negative_slope=0.0, dropout=0.0, text_dim=c["text_dim"], patch_dim=c["patch_dim"])
    torch.manual_seed(seed)
    if variant == "vanilla":
        m = importlib.import_module("vanilla_gan_unconditional")
        t = m.WGAN_GP_nocond(input_dims=G, latent_dims=c["latent"], vocab_sizes=[], generator_dims=[H, H, G],
                             discriminator_dims=[H, H, 1], optimizer=optimizer, **kw)
        t.build_WGAN_GP_nocond()
    else:
        m = importlib.import_module({"paper": "conditional_gan_cross_attention_with_film",
                                     "film": "conditional_gan_film"}[variant])
        t = m.WGAN_GP(input_dims=G, latent_dims=c["latent"], embedding_dims=c["embed"], generator_dims=[H, H, G],
                      discriminator_dims=[H, H, 1], text_embedding_dims=c["text_dim"],
                      patches_embedding_dims=c["patch_dim"], optimizer=optimizer, **kw)
        t.build_WGAN_GP()
    t.init_train()
    for a, b in ((o.gen, t.gen), (o.disc, t.disc)):
        for (k1, v1), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), k1
    return o, t


def batch(variant, B, seed):
    c = SMALL
    return restated.synthetic_batch(variant, B, c["G"], P=c["P"], T=c["T"], seed=seed, ragged=True,
                                    text_dim=c["text_dim"], patch_dim=c["patch_dim"])


def call_train(t, variant, x, cond, zs=None, alphas=None):
    """train() in each variant's own argument order (…with_film.py:463, conditional_gan_film.py:432,
    vanilla_gan_unconditional.py:417); cond is in model-argument order."""
    if variant == "vanilla":
        t.train(x, zs=zs, alphas=alphas)
    elif variant == "paper":
        patches, ppad, text, tpad = cond
        t.train(x, text, tpad, patches, ppad, zs=zs, alphas=alphas)
    else:
        text, patches, ppad = cond
        t.train(x, text, patches, ppad, zs=zs, alphas=alphas)


def noise(B, n_critic=5, seed=5):
    g = torch.Generator().manual_seed(seed)
    zs = [torch.randn(B, SMALL["latent"], generator=g) for _ in range(n_critic + 1)]
    alphas = [torch.rand(B, 1, generator=g) for _ in range(n_critic)]
    return zs, alphas


def snapshot(tr):
    return ({k: v.clone() for k, v in tr.disc.state_dict().items()}, {k: v.clone() for k, v in tr.gen.state_dict().items()})


def update_cosine(before, ref_net, net):
    """Cosine between the two UPDATE vectors (weights after - weights before, every tensor concatenated). The first
    Adam / RMSprop steps are lr * sign(g) / 10 lr * sign(g) for every entry, so rounding in near-zero gradient entries
    becomes full steps: the cosine is about the share of entries whose sign agrees (the measure
    tests/test_gpu_parity_full.py uses on the B200)."""
    ur = torch.cat([(v - before[k]).flatten() for k, v in ref_net.state_dict().items()])
    ug = torch.cat([(v - before[k]).flatten() for k, v in net.state_dict().items()])
    return float(ur @ ug / (ur.norm() * ug.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("variant,optimizer", [("vanilla", "adam"), ("vanilla", "rms_prop"), ("paper", "adam"),
                                               ("film", "rms_prop")])
def test_train_call_matches_the_oracle(host, variant, optimizer):
    """One train() (5 critic steps + 1 generator step) through the reference-facing method: loss statistics and the
    weights both networks end up with."""
    o, t = make(variant, optimizer)
    B = 8
    x, cond = batch(variant, B, seed=3)
    zs, alphas = noise(B)
    bd, bg = snapshot(o)
    o.train(x, cond, zs, alphas)
    call_train(t, variant, x, cond, zs, alphas)
    scale = max(1.0, float(np.abs(o.d_batch_loss).max()))
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * scale, (t.d_batch_loss, o.d_batch_loss)
    assert abs(float(t.g_batch_loss[0]) - float(o.g_batch_loss[0])) <= TOL * max(1.0, abs(float(o.g_batch_loss[0])))
    cd, cg = update_cosine(bd, o.disc, t.disc), update_cosine(bg, o.gen, t.gen)
    print(f"{variant}/{optimizer}: update cosine critic {cd:.4f} generator {cg:.4f}")
    assert cd > COS_FLOOR and cg > COS_FLOOR, (cd, cg)
    # observable side effects of the reference's train_gen (:433-438)
    assert all(not p.requires_grad for p in t.disc.parameters()) and all(p.requires_grad for p in t.gen.parameters())


def test_noise_drawn_inside_train_follows_the_reference_order(host):
    """Without explicit noise train() draws z, alpha, z, alpha, ..., z from torch's global stream — the order of the
    reference's loop (:463-477, :354) — so seeding both sides gives the same call."""
    o, t = make("vanilla", "adam")
    B = 8
    x, cond = batch("vanilla", B, seed=3)
    bd, bg = snapshot(o)
    torch.manual_seed(123)
    o.train(x, cond)
    torch.manual_seed(123)
    call_train(t, "vanilla", x, cond)
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * max(1.0, float(np.abs(o.d_batch_loss).max()))
    assert update_cosine(bd, o.disc, t.disc) > COS_FLOOR and update_cosine(bg, o.gen, t.gen) > COS_FLOOR


def test_fit_rebuilds_the_networks_and_trains_the_new_ones(host, tmp_path):
    """fit() builds the networks again (reference :620-623). A trainer that was already built and stepped must then
    train — and checkpoint — the NEW pair, not the one its flat buffers were made from."""
    _, t = make("vanilla", "adam", results_dire=str(tmp_path))
    x, cond = batch("vanilla", 8, seed=3)
    call_train(t, "vanilla", x, cond)
    old_gen = t.gen
    data = [(batch("vanilla", 8, seed=20 + i)[0],) for i in range(3)]
    t.fit(data, epochs=1)
    assert t.gen is not old_gen and t._flat_gen.module is t.gen and t._flat_disc.module is t.disc
    w0 = t.gen.final_layer.weight.detach().clone()
    call_train(t, "vanilla", x, cond)
    assert float((t.gen.final_layer.weight - w0).abs().max()) > 0.0
    saved = torch.load(tmp_path / "generator_last_epoch.pt")
    assert torch.equal(saved["final_layer.weight"], w0)           # what fit() trained is what it saved
    assert len(t.loss_dict["d loss"]) == 1 and np.isfinite(t.loss_dict["g loss"][0])


@pytest.mark.parametrize("variant", ["vanilla", "paper"])
def test_engines_of_other_batch_sizes_follow_the_weights(host, variant):
    """len(dataset) % B != 0 and a validation loader of another batch size: every batch size has its own engine with
    its own bf16 weight shadows. Training with one engine must be seen by the others — compare with the oracle taking
    the same sequence of batches, and generation at a third batch size with a fresh trainer loaded from the weights."""
    o, t = make(variant, "adam")
    sizes = [8, 3, 8]                            # 11 rows at B = 8, then the next epoch
    bd, bg = snapshot(o)
    for i, B in enumerate(sizes):
        x, cond = batch(variant, B, seed=40 + i)
        zs, alphas = noise(B, seed=60 + i)
        o.train(x, cond, zs, alphas)
        call_train(t, variant, x, cond, zs, alphas)
    assert len(t._engines) == 2
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * max(1.0, float(np.abs(o.d_batch_loss).max()))
    cd, cg = update_cosine(bd, o.disc, t.disc), update_cosine(bg, o.gen, t.gen)
    print(f"{variant}: update cosine over 18 steps on two engines: critic {cd:.4f} generator {cg:.4f}")
    assert cd > COS_FLOOR and cg > COS_FLOOR, (cd, cg)
    # a validation batch of 5 rows, first on a new engine, then again after more training on the B = 8 engine
    xv, cv = batch(variant, 5, seed=77)

    def generate(tr):
        torch.manual_seed(9)
        if variant == "vanilla":
            return tr.generate_samples(xv)[1]
        patches, ppad, text, tpad = cv
        return tr.generate_samples(xv, text, tpad, patches, ppad)[1]

    def fresh_copy():
        _, f = make(variant, "adam", seed=99)
        f.gen.load_state_dict(t.gen.state_dict())
        f.disc.load_state_dict(t.disc.state_dict())
        return f

    assert torch.equal(generate(t), generate(fresh_copy()))
    x, cond = batch(variant, 8, seed=50)
    call_train(t, variant, x, cond, *noise(8, seed=70))
    assert len(t._engines) == 3
    assert torch.equal(generate(t), generate(fresh_copy()))      # the B = 5 engine refreshed its shadows


def test_lr_halving_through_param_groups_reaches_the_kernel(host):
    """fit() halves both learning rates through optimizer.param_groups (:649-657): the engine's update must use them."""
    o, t = make("vanilla", "rms_prop")
    B = 8
    x, cond = batch("vanilla", B, seed=3)
    for tr in (o, t):
        for opt in (tr.optimizer_disc, tr.optimizer_gen):
            for g in opt.param_groups:
                g["lr"] *= 0.5
    w0 = t.gen.final_layer.weight.detach().clone()
    zs, alphas = noise(B)
    bd, bg = snapshot(o)
    o.train(x, cond, zs, alphas)
    call_train(t, "vanilla", x, cond, zs, alphas)
    assert update_cosine(bd, o.disc, t.disc) > COS_FLOOR and update_cosine(bg, o.gen, t.gen) > COS_FLOOR
    ud = torch.cat([(v - bd[k]).flatten() for k, v in t.disc.state_dict().items()])
    ur = torch.cat([(v - bd[k]).flatten() for k, v in o.disc.state_dict().items()])
    assert 0.9 < float(ud.norm() / ur.norm()) < 1.1               # half the rate moved the weights half as far
    # RMSprop's first step moves every weight with a non-zero gradient by lr / sqrt(1 - alpha) = 10 lr = 2.5e-3
    step = float((t.gen.final_layer.weight - w0).abs().max())
    assert 2.0e-3 <= step <= 2.6e-3, step
    t._epoch_lr_decay(100, 100)
    assert t.optimizer_gen.param_groups[0]["lr"] == pytest.approx(1.25e-4)
    t._epoch_lr_decay(101, 100)
    assert t.optimizer_gen.param_groups[0]["lr"] == pytest.approx(1.25e-4)


@pytest.mark.parametrize("optimizer", ["adam", "rms_prop"])
def test_checkpoint_of_networks_and_optimizers_resumes_bitwise(host, optimizer):
    """torch.save of the four state_dicts after some training, loaded into a new trainer: both continue identically
    (the optimizer state lives in the flat buffers the kernel updates; state_dict() sees it through views)."""
    _, t = make("vanilla", optimizer)
    B = 8
    x, cond = batch("vanilla", B, seed=3)
    call_train(t, "vanilla", x, cond, *noise(B, seed=1))
    buf = io.BytesIO()
    torch.save({"gen": t.gen.state_dict(), "disc": t.disc.state_dict(), "og": t.optimizer_gen.state_dict(),
                "od": t.optimizer_disc.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf)
    _, r = make("vanilla", optimizer, seed=99)
    call_train(r, "vanilla", x, cond, *noise(B, seed=2))           # r has engines and state of its own by now
    r.gen.load_state_dict(ck["gen"])
    r.disc.load_state_dict(ck["disc"])
    r.optimizer_gen.load_state_dict(ck["og"])
    r.optimizer_disc.load_state_dict(ck["od"])
    x2, cond2 = batch("vanilla", B, seed=4)
    call_train(t, "vanilla", x2, cond2, *noise(B, seed=3))
    call_train(r, "vanilla", x2, cond2, *noise(B, seed=3))
    for (k, a), (_, b) in zip(t.gen.state_dict().items(), r.gen.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(t.disc.state_dict().items(), r.disc.state_dict().items()):
        assert torch.equal(a, b), k
    assert np.array_equal(t.d_batch_loss, r.d_batch_loss)


def test_single_steps_and_module_calls_through_the_trainer(host):
    """train_disc / train_gen / gradient_penalty / gen(...) / disc(...) in the reference's signatures (:351-461)."""
    o, t = make("paper", "adam")
    B = 8
    x, cond = batch("paper", B, seed=3)
    patches, ppad, text, tpad = cond
    zs, alphas = noise(B)
    o.train_disc(x, zs[0], cond, alphas[0])
    t.train_disc(x, zs[0], text, tpad, patches, ppad, alpha=alphas[0])
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * max(1.0, float(np.abs(o.d_batch_loss).max()))
    assert all(p.requires_grad for p in t.disc.parameters()) and all(not p.requires_grad for p in t.gen.parameters())
    # the prototype encoder layer the reference registers but never runs keeps grad None (:114)
    assert all(p.grad is None for p in t.disc.patches_transformer_layer.parameters())
    assert t.disc.final_layer.weight.grad is not None
    o.train_gen(zs[1], cond)
    t.train_gen(zs[1], text, tpad, patches, ppad)
    assert abs(float(t.g_batch_loss[0]) - float(o.g_batch_loss[0])) <= TOL * max(1.0, abs(float(o.g_batch_loss[0])))
    t.gen.eval(), t.disc.eval(), o.gen.eval(), o.disc.eval()
    with torch.no_grad():
        fake_o = o.gen(zs[2], *cond)
        fake_t = t.gen(zs[2], *cond)
        assert float((fake_t - fake_o).abs().max()) <= TOL * float(fake_o.abs().max())
        s_o, s_t = o.disc(x, *cond), t.disc(x, *cond)
        assert s_t.shape == s_o.shape == (B, 1)
        assert float((s_t - s_o).abs().max()) <= TOL * max(1.0, float(s_o.abs().max()))
    t.disc.train(), o.disc.train()
    gp_o = o.gradient_penalty(x, fake_o, cond, alphas[1])
    gp_t = t.gradient_penalty(x, fake_o, patches, ppad, text, tpad, alpha=alphas[1])
    assert abs(float(gp_t) - float(gp_o)) <= TOL * max(1.0, abs(float(gp_o)))


# ----------------------------------------------------------------------- generate_samples_all, class-balanced branch
class _Cases(torch.utils.data.Dataset):
    """Film-layout items (text, genes, patches, pad, disease type, primary site) with the loaders' random patch subset
    (np.random.choice per read of an over-long case, multi_patch_gan_dataloader.py:33-36) and skewed class counts."""

    def __init__(self, n=45, P=4):
        g = torch.Generator().manual_seed(3)
        self.text = torch.randn(n, SMALL["text_dim"], generator=g)
        self.genes = torch.randn(n, SMALL["G"], generator=g)
        self.patches = [torch.randn(2 + i % 5, SMALL["patch_dim"], generator=g) for i in range(n)]   # 2..6 patches
        self.disease = torch.tensor([0 if i % 3 else (1 if i % 9 else 2) for i in range(n)])          # 30 / 10 / 5
        self.site = torch.arange(n) % 4
        self.P = P

    def __len__(self):
        return len(self.genes)

    def __getitem__(self, i):
        p = self.patches[i]
        if p.shape[0] > self.P:
            p = p[np.random.choice(p.shape[0], self.P, replace=False)]
        else:
            p = torch.cat((p, torch.zeros(self.P - p.shape[0], p.shape[1])))
        return self.text[i], self.genes[i], p, torch.zeros(self.P, dtype=torch.bool), self.disease[i], self.site[i]


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_class_balanced_generation_matches_the_reference(host, tmp_path):
    """conditional_gan_concat.py:453-552 with balanced=True against the drop-in on the same weights, numpy and torch
    seeds: same real rows, same generated labels in the same (shuffled) order — i.e. the same np.random stream,
    including the per-field dataset reads — and the generated profiles within the bf16 tolerance."""
    import sys

    import conditional_gan_concat as mine_mod

    H, G = SMALL["hidden"], SMALL["G"]
    kw = dict(input_dims=G, latent_dims=SMALL["latent"], input_embedding_dims=SMALL["text_dim"], embedding_dims=SMALL["embed"],
              generator_dims=[H, H, G], discriminator_dims=[H, H, 1], condition_on="text",
              results_dire=str(tmp_path))        # (the reference's constructor creates the directory, :260)
    torch.manual_seed(5)
    mine = mine_mod.WGAN_GP(**kw)
    mine.build_WGAN_GP()
    sys.modules.pop("conditional_gan_concat", None)
    ref_mod = ref_shim.load("conditional_gan_concat")
    assert ref_mod.__file__ != mine_mod.__file__
    ref = ref_mod.WGAN_GP(**kw)
    ref.build_WGAN_GP()
    ref.gen.load_state_dict(mine.gen.state_dict())
    loader = torch.utils.data.DataLoader(_Cases(), batch_size=8, shuffle=False)
    outs = []
    for tr in (ref, mine):
        np.random.seed(11)
        torch.manual_seed(12)
        outs.append(tr.generate_samples_all(loader, num_repeats=2, balanced=True, balanced_max_oversample=3))
    r, m = outs
    assert len(r) == len(m) == 4
    # 30 rows of class 0; class 1: 10 + min(20, 30) = 30; class 2: 5 + min(25, 15) = 20; two repeats
    assert m[1].shape == r[1].shape == (2 * (30 + 30 + 20), G)
    assert np.array_equal(r[0], m[0]) and np.array_equal(r[2], m[2]) and np.array_equal(r[3], m[3])
    assert np.abs(m[1] - r[1]).max() <= TOL * np.abs(r[1]).max()
    assert sorted(mine._engines) == [64]                       # chunks of 64, the remainders padded to it
    # the unbalanced 4-tuple as well (partial last batch: 45 = 5 x 8 + 5)
    outs = []
    for tr in (ref, mine):
        torch.manual_seed(13)
        np.random.seed(14)
        outs.append(tr.generate_samples_all(loader))
    r, m = outs
    assert len(m) == 4 and np.array_equal(r[0], m[0]) and np.array_equal(r[2], m[2]) and np.array_equal(r[3], m[3])
    assert np.abs(m[1] - r[1]).max() <= TOL * np.abs(r[1]).max()
    sys.modules.pop("conditional_gan_concat", None)


def test_generated_array_tuples_follow_each_script(host, tmp_path):
    """6 arrays from the film / img scripts (balanced=True raises, as the reference's NameError does), 4 from concat /
    attn; save_generated_arrays writes twelve or eight files."""
    from gemmgan_b200.trainer import save_generated_arrays

    H, G = SMALL["hidden"], SMALL["G"]
    loader = torch.utils.data.DataLoader(_Cases(n=11), batch_size=8, shuffle=False)
    for name, n_out in (("conditional_gan_film", 6), ("conditional_gan_attention", 4)):
        m = importlib.import_module(name)
        t = m.WGAN_GP(input_dims=G, latent_dims=SMALL["latent"], embedding_dims=SMALL["embed"], generator_dims=[H, H, G],
                      discriminator_dims=[H, H, 1], text_embedding_dims=SMALL["text_dim"],
                      patches_embedding_dims=SMALL["patch_dim"])
        t.build_WGAN_GP()
        out = t.generate_samples_all(loader)
        assert len(out) == n_out and out[1].shape == (11, G) and np.array_equal(out[2], out[3])
        folder = tmp_path / name
        save_generated_arrays(str(folder), out, out)
        assert len(list(folder.iterdir())) == 2 * n_out
        if n_out == 6:
            assert np.array_equal(np.load(folder / "test_primary_site_real.npy"), np.arange(11) % 4)
            with pytest.raises(NotImplementedError):
                t.generate_samples_all(loader, balanced=True)
        else:
            balanced = t.generate_samples_all(loader, balanced=True)
            assert len(balanced) == 4 and balanced[1].shape[0] == len(balanced[3]) > 11


@pytest.mark.parametrize("variant,B", [("paper", 1), ("vanilla", 1), ("film", 2)])
def test_smallest_batches_train(host, variant, B):
    """A loader without drop_last can end an epoch with a batch of one or two rows: the step must still be the
    reference's (batch means over B rows, per-row gradient penalty)."""
    o, t = make(variant, "adam")
    x, cond = batch(variant, B, seed=3)
    zs, alphas = noise(B)
    bd, bg = snapshot(o)
    o.train(x, cond, zs, alphas)
    call_train(t, variant, x, cond, zs, alphas)
    scale = max(1.0, float(np.abs(o.d_batch_loss).max()))
    assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * scale, (t.d_batch_loss, o.d_batch_loss)
    assert abs(float(t.g_batch_loss[0]) - float(o.g_batch_loss[0])) <= TOL * max(1.0, abs(float(o.g_batch_loss[0])))
    assert update_cosine(bd, o.disc, t.disc) > COS_FLOOR and update_cosine(bg, o.gen, t.gen) > COS_FLOOR


def test_single_row_batch_is_refused_by_the_batchnorm_script(host):
    """conditional_gan_attention.py: nn.BatchNorm1d in the generator raises on a training batch of one row; so does the
    drop-in, with torch's message (eval-mode generation of one row works, as in the reference)."""
    m = importlib.import_module("conditional_gan_attention")
    H, G = SMALL["hidden"], SMALL["G"]
    t = m.WGAN_GP(input_dims=G, latent_dims=SMALL["latent"], embedding_dims=SMALL["embed"], generator_dims=[H, H, G],
                  discriminator_dims=[H, H, 1], text_embedding_dims=SMALL["text_dim"],
                  patches_embedding_dims=SMALL["patch_dim"])
    t.build_WGAN_GP()
    t.init_train()
    x, (text, patches, ppad) = batch("film", 1, seed=3)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel when training"):
        t.train(x, text, patches, ppad)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        t.train_gen(torch.randn(1, SMALL["latent"]), text, patches, ppad)
    real, fake = t.generate_samples(x, text, patches, ppad)
    assert fake.shape == (1, G) and torch.isfinite(fake).all()


def test_wrong_tensor_sizes_fail_on_the_host(host):
    """The library reads raw pointers with the engine's static sizes: a batch tensor of another width must raise with
    the expected shape (the reference raises a matmul shape error), never be read out of bounds."""
    _, t = make("paper", "adam")
    x, (patches, ppad, text, tpad) = batch("paper", 8, seed=3)
    t.train(x, text, tpad, patches, ppad)
    with pytest.raises(ValueError, match=r"gene expression: expected shape \(8, 203\)"):
        t.train(x[:, :200], text, tpad, patches, ppad)
    with pytest.raises(ValueError, match="patches: expected shape"):
        t.train(x, text, tpad, patches[:, :, :24], ppad)
    with pytest.raises(ValueError, match="text embedding: expected shape"):
        t.train(x, text[:, :, :16], tpad, patches, ppad)
    with pytest.raises(ValueError, match="patch padding mask: expected shape"):
        t.train(x, text, tpad, patches, ppad[:4])
    with pytest.raises(ValueError, match="z: expected shape"):
        t.gen(torch.randn(8, SMALL["latent"] + 8), patches, ppad, text, tpad)
    with pytest.raises(ValueError, match="fake data: expected shape"):
        t.gradient_penalty(x, x[:, :100], patches, ppad, text, tpad)
    t.train(x, text, tpad, patches, ppad)           # and the trainer is still usable
    assert np.isfinite(t.d_batch_loss).all()


def test_critic_steps_after_generation_see_an_eval_mode_generator(host):
    """The reference's train_disc never calls gen.train() (conditional_gan_attention.py:322; …with_film.py:390): after
    generate_samples (gen.eval()) the critic steps of the next train() run the generator in eval mode — in the
    BatchNorm script on the running statistics, which then do not move — until train_gen switches it back. Checked on
    the BatchNorm variant, where the difference is deterministic, against the oracle (which keeps the reference's
    mode handling) and against the same call with the generator in training mode."""
    m = importlib.import_module("conditional_gan_attention")
    H, G = SMALL["hidden"], SMALL["G"]
    kw = dict(input_dims=G, latent_dims=SMALL["latent"], embedding_dims=SMALL["embed"], generator_dims=[H, H, G],
              discriminator_dims=[H, H, 1], text_embedding_dims=SMALL["text_dim"], patches_embedding_dims=SMALL["patch_dim"],
              optimizer="adam")
    B = 8
    x, cond = batch("film", B, seed=3)
    text, patches, ppad = cond
    zs, alphas = noise(B)
    results = {}
    for mode in ("eval", "train"):
        torch.manual_seed(11)
        o = restated.OracleWGANGP("attn", G, latent=SMALL["latent"], embed=SMALL["embed"], hidden=H, optimizer="adam",
                                  negative_slope=0.0, dropout=0.0, text_dim=SMALL["text_dim"], patch_dim=SMALL["patch_dim"])
        torch.manual_seed(11)
        t = m.WGAN_GP(**kw)
        t.build_WGAN_GP()
        t.init_train()
        for tr in (o, t):                  # one ordinary call first: the running statistics are no longer (0, 1)
            if tr is o:
                tr.train(x, cond, *noise(B, seed=9))
            else:
                tr.train(x, text, patches, ppad, *noise(B, seed=9))
        if mode == "eval":
            o.gen.eval()
            t.generate_samples(x, text, patches, ppad)      # leaves the generator in eval mode (:603)
            assert not t.gen.training
        tracked = int(t.gen.attn_bn.num_batches_tracked)
        bd, bg = snapshot(o)
        o.train(x, cond, zs, alphas)
        t.train(x, text, patches, ppad, zs=zs, alphas=alphas)
        assert t.gen.training and o.gen.training            # train_gen switched it back (:427)
        scale = max(1.0, float(np.abs(o.d_batch_loss).max()))
        assert np.abs(t.d_batch_loss - o.d_batch_loss).max() <= TOL * scale, (mode, t.d_batch_loss, o.d_batch_loss)
        assert update_cosine(bd, o.disc, t.disc) > COS_FLOOR
        ob, tb = o.gen.attn_bn, t.gen.attn_bn
        assert float((tb.running_mean - ob.running_mean).abs().max()) <= TOL * max(1.0, float(ob.running_mean.abs().max()))
        assert float((tb.running_var - ob.running_var).abs().max()) <= TOL * max(1.0, float(ob.running_var.abs().max()))
        # eval mode: only train_gen's forward is a training-mode batch; training mode: all six
        want = 1 if mode == "eval" else 6
        assert int(tb.num_batches_tracked) - tracked == want and int(ob.num_batches_tracked) - tracked == want
        results[mode] = (t.d_batch_loss.copy(), tb.running_mean.clone())
    # and the two modes do differ (one momentum update of the statistics instead of six)
    assert not torch.allclose(results["eval"][1], results["train"][1], rtol=1e-3, atol=1e-5)
