"""fit() of EVERY drop-in trainer on the CPU suite: the outer loop of the reference scripts — build, init_train, the
loader's batch tuples in each variant's own layout and argument order, a last partial batch, per-epoch loss_dict
entries, checkpoints, generate_samples_all — on the host-emulated engine (tests/host_trainer.py), against the oracle
(oracle/restated.py, pinned to the unmodified reference) stepping through the same loader with the same torch seed.

Reference loops: src/conditional_gan_cross_attention_with_film.py:619-744, conditional_gan_film.py:592-700,
conditional_gan_concat.py:597-705, conditional_gan_img_transformer.py, conditional_gan_attention.py:523-600,
benchmark_generative_model.py:559-640, vanilla_gan_unconditional.py:520-615. Batch layouts: gemmgan_b200/synthetic.py.
The GPU suite runs the same loops on the B200 (tests/test_gpu_fit.py)."""
import importlib
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

import emu_build
import host_trainer
from gemmgan_b200 import _abi_decl as A
from gemmgan_b200.synthetic import synthetic_tensors
from oracle import restated

G, B, N = 203, 8, 11          # 11 rows at B = 8: batches of 8 and 3 (the reference loaders have no drop_last)
E, H, LZ, DT, DP, P, T = 32, 32, 16, 24, 32, 5, 3
MODULES = {"paper": "conditional_gan_cross_attention_with_film", "cross": "conditional_gan_cross_attention",
           "film": "conditional_gan_film", "img": "conditional_gan_img_transformer", "attn": "conditional_gan_attention",
           "concat": "conditional_gan_concat", "concat_image": "conditional_gan_concat",
           "label": "benchmark_generative_model", "vanilla": "vanilla_gan_unconditional"}


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("engine", tmp_path_factory.mktemp("cuda_emu"), cudart=True)
    A.declare(L)
    return L


@pytest.fixture()
def host(emu, monkeypatch):
    return host_trainer.apply(monkeypatch.setattr, emu)


def layout_of(variant):
    return {"paper": "paper", "cross": "paper", "vanilla": "vanilla", "label": "label"}.get(variant, "film")


def loader_for(variant, seed=1):
    tensors = synthetic_tensors(layout_of(variant), N, G, P, T, text_dim=DT, patch_dim=DP, seed=seed, ragged=True)
    if variant == "concat_image":      # one encoder over patch-sized vectors: the text slot is unused
        pass
    return DataLoader(TensorDataset(*tensors), batch_size=B, shuffle=False)


def make(variant, results_dire):
    m = importlib.import_module(MODULES[variant])
    kw = dict(input_dims=G, latent_dims=LZ, generator_dims=[H, H, G], discriminator_dims=[H, H, 1], optimizer="adam",
              results_dire=results_dire, freq_print=100)
    if variant == "vanilla":
        return m.WGAN_GP_nocond(vocab_sizes=[], **kw)
    if variant == "label":
        return m.WGAN_GP_benchmark(vocab_sizes=[10, 10], freq_compute_test=2, **kw)
    if variant.startswith("concat"):
        image = variant == "concat_image"
        return m.WGAN_GP(embedding_dims=E, input_embedding_dims=DP if image else DT,
                         condition_on="image" if image else "text", **kw)
    return m.WGAN_GP(embedding_dims=E, text_embedding_dims=DT, patches_embedding_dims=DP, **kw)


def oracle_for(variant):
    return restated.OracleWGANGP(variant, G, latent=LZ, embed=E, hidden=H, optimizer="adam", negative_slope=0.0,
                                 dropout=0.0, text_dim=DT, patch_dim=DP)


def split(variant, batch):
    """(genes, cond in model-argument order) of one loader tuple."""
    lay = layout_of(variant)
    if lay == "vanilla":
        return batch[0], ()
    if lay == "label":
        return batch[0], (batch[1], batch[2])
    if lay == "paper":
        text, tpad, genes, patches, ppad = batch[:5]
        return genes, (patches, ppad, text, tpad)
    text, genes, patches, ppad = batch[:4]
    return genes, (text, patches, ppad)


def oracle_fit(o, variant, loader, epochs):
    """The reference's loop on the oracle: per-epoch MEAN of d_batch_loss and SUM of g_batch_loss over the batches —
    every script divides only the critic's (…with_film.py:695-699, vanilla_gan_unconditional.py:602-606)."""
    hist = {"d loss": [], "d real loss": [], "d fake loss": [], "g loss": []}
    for _ in range(epochs):
        d_sum, g_sum, n = 0.0, 0.0, 0
        for batch in loader:
            x, cond = split(variant, batch)
            o.train(x, cond)
            d_sum, g_sum, n = d_sum + o.d_batch_loss, g_sum + o.g_batch_loss, n + 1
        d = d_sum / n
        hist["d loss"].append(d[0]), hist["d real loss"].append(d[1]), hist["d fake loss"].append(d[2])
        hist["g loss"].append(g_sum[0])
    return hist


# epochs between two halvings of both learning rates: `if epoch % N == 0 and epoch != 0` in each script's fit()
# (…with_film.py:649, conditional_gan_cross_attention.py:619, conditional_gan_film.py:614, conditional_gan_img_transformer.py:597,
# conditional_gan_attention.py:547, conditional_gan_concat.py:605, vanilla_gan_unconditional.py:558, benchmark_generative_model.py:585)
LR_HALVING = {"paper": 100, "cross": 100, "film": 100, "img": 100, "attn": 50, "concat": 50, "concat_image": 50,
              "vanilla": 50, "label": 50}


def test_lr_halving_table_is_the_reference():
    import re

    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree only exists in the build container")
    for variant, module in MODULES.items():
        src = open(os.path.join(ref_shim.REF_SRC, module + ".py")).read()
        found = re.findall(r"if epoch % (\d+) == 0 and epoch != 0:", src)
        assert found == [str(LR_HALVING[variant])], (module, found)


@pytest.mark.parametrize("variant", sorted(MODULES))
def test_fit_follows_the_reference_loop(host, variant, tmp_path, monkeypatch):
    calls = []
    decay = host.TrainerBase._epoch_lr_decay
    monkeypatch.setattr(host.TrainerBase, "_epoch_lr_decay",
                        lambda self, epoch, every: (calls.append(every), decay(self, epoch, every))[1])
    epochs = 2 if variant in ("vanilla", "label", "concat", "paper") else 1      # (suite time)
    loader = loader_for(variant)
    torch.manual_seed(7)
    o = oracle_for(variant)
    ref = oracle_fit(o, variant, loader, epochs)
    torch.manual_seed(7)
    t = make(variant, str(tmp_path))
    if variant in ("attn", "label", "vanilla"):   # fit(train_data, test_data, epochs, val) in these three scripts
        t.fit(loader, None, epochs)
    else:
        t.fit(loader, None, None, epochs=epochs)
    # same initial weights were drawn (fit() builds the nets from the seeded stream, as the reference does)
    for k in ref:
        got, want = np.asarray(t.loss_dict[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
        assert got.shape == (epochs,) and np.isfinite(got).all(), (k, got)
        tol = 3e-2 * max(1.0, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= tol, (variant, k, got, want)
    assert calls == [LR_HALVING[variant]] * epochs
    assert len(t._engines) == 2                 # B = 8 and the last partial batch of 3
    assert os.path.exists(tmp_path / "generator_last_epoch.pt") and os.path.exists(tmp_path / "discriminator_last_epoch.pt")
    sd = torch.load(tmp_path / "generator_last_epoch.pt")
    assert list(sd) == list(o.gen.state_dict()) and all(torch.equal(sd[k], v) for k, v in t.gen.state_dict().items())
    # generation over the whole loader (partial batch included), eval mode
    out = t.generate_samples_all(loader)
    real, fake = out[0], out[1]
    assert real.shape == fake.shape == (N, G) and np.isfinite(fake).all()
    assert np.array_equal(real, loader.dataset.tensors[{"paper": 2, "film": 1}.get(layout_of(variant), 0)].numpy())
    assert not t.gen.training
