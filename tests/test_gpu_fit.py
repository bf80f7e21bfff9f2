"""GPU tests of the drop-in trainers' outer loop: fit() over a DataLoader of the reference's batch tuples (every
variant), checkpoint files with the reference's names, generate_samples_all, and the next-batch prefetch of
train(prefetch=...) (same numbers with and without it)."""
import importlib
import os

import numpy as np
import pytest
import torch

from gemmgan_b200.synthetic import synthetic_loader, synthetic_tensors

pytestmark = pytest.mark.gpu

G, B = 300, 16
DIMS = dict(latent_dims=32, embedding_dims=32, generator_dims=[32, 32, G], discriminator_dims=[32, 32, 1])
SMALLF = dict(text_embedding_dims=24, patches_embedding_dims=40)


def _make(variant, tmp):
    kw = dict(input_dims=G, optimizer="adam", results_dire=str(tmp), **DIMS)
    if variant == "vanilla":
        m = importlib.import_module("vanilla_gan_unconditional")
        kw.pop("embedding_dims")
        return m.WGAN_GP_nocond(vocab_sizes=[], **kw), "vanilla"
    if variant.startswith("concat"):
        m = importlib.import_module("conditional_gan_concat")
        image = variant == "concat_image"
        return m.WGAN_GP(input_embedding_dims=40 if image else 24, condition_on="image" if image else "text", **kw), "film"
    mod = {"paper": "conditional_gan_cross_attention_with_film", "cross": "conditional_gan_cross_attention",
           "film": "conditional_gan_film", "img": "conditional_gan_img_transformer",
           "attn": "conditional_gan_attention"}[variant]
    m = importlib.import_module(mod)
    return m.WGAN_GP(**kw, **SMALLF), ("paper" if variant in ("paper", "cross") else "film")


@pytest.mark.parametrize("variant", ["paper", "cross", "film", "img", "concat", "concat_image", "vanilla", "attn"])
def test_fit_runs_and_saves_checkpoints(variant, tmp_path):
    torch.manual_seed(0)
    t, layout = _make(variant, tmp_path)
    loader = synthetic_loader(layout, n_samples=3 * B, batch_size=B, n_genes=G, n_patches=5, n_tokens=3, seed=1,
                              text_dim=24, patch_dim=40, ragged=True)
    if variant == "attn":   # conditional_gan_attention.py:523: fit(train_data, test_data, epochs, val)
        t.fit(loader, None, epochs=2)
        bn = t.gen.attn_bn   # 2 epochs x 3 batches x (5 critic + 1 generator) training-mode generator forwards
        assert int(bn.num_batches_tracked) == 36 and bn.running_var.min().item() > 0
        assert not torch.equal(bn.running_mean, torch.zeros_like(bn.running_mean))
    elif variant == "vanilla":   # vanilla_gan_unconditional.py:543: fit(train_data, test_data, epochs, val)
        t.fit(loader, None, 2)
    else:
        t.fit(loader, None, None, epochs=2)
    for k in ("d loss", "d real loss", "d fake loss", "g loss"):
        assert len(t.loss_dict[k]) == 2 and np.isfinite(t.loss_dict[k]).all(), (k, t.loss_dict[k])
    # the reference's checkpoint names (conditional_gan_cross_attention_with_film.py:743-744, vanilla :614-615)
    assert os.path.exists(tmp_path / "generator_last_epoch.pt") and os.path.exists(tmp_path / "discriminator_last_epoch.pt")
    sd = torch.load(tmp_path / "generator_last_epoch.pt")
    assert list(sd.keys()) == list(t.gen.state_dict().keys())
    if variant != "vanilla":
        out = t.generate_samples_all(loader)
        assert out[0].shape == (3 * B, G) and out[1].shape == (3 * B, G) and np.isfinite(out[1]).all()


def test_label_conditioned_baseline_fit(tmp_path):
    """benchmark_generative_model.WGAN_GP_benchmark.fit(train, test, epochs) over (genes, disease type, primary site)
    batches; checkpoints only at the last epoch when it is a freq_compute_test multiple (reference :643-656)."""
    m = importlib.import_module("benchmark_generative_model")
    torch.manual_seed(0)
    t = m.WGAN_GP_benchmark(input_dims=G, latent_dims=32, vocab_sizes=[10, 10], generator_dims=[32, 32, G],
                            discriminator_dims=[32, 32, 1], optimizer="rms_prop", freq_compute_test=1,
                            results_dire=str(tmp_path))
    loader = synthetic_loader("label", n_samples=3 * B, batch_size=B, n_genes=G, seed=1)
    t.fit(loader, None, epochs=2)
    for k in ("d loss", "d real loss", "d fake loss", "g loss"):
        assert len(t.loss_dict[k]) == 2 and np.isfinite(t.loss_dict[k]).all(), (k, t.loss_dict[k])
    sd = torch.load(tmp_path / "generator_last_epoch.pt")
    assert list(sd.keys()) == list(t.gen.state_dict().keys()) and "categorical_embedding.1.weight" in sd
    real, gen, cats, _, sites, _ = t.generate_samples_all(loader)
    assert real.shape == (3 * B, G) and gen.shape == (3 * B, G) and np.isfinite(gen).all()
    assert len(cats) == 3 * B and len(sites) == 3 * B
    # the embedding rows of labels that occur in the data have moved, the others have not
    e0 = t.gen.categorical_embedding[0].weight.detach().cpu()
    torch.manual_seed(0)
    fresh, _ = m.WGAN_GP_model_benchmark(32, G, [], [10, 10], [32, 32, G], [32, 32, 1])
    seen = torch.zeros(10, dtype=torch.bool)
    for b in loader:
        seen[b[1]] = True
    moved = (e0 - fresh.categorical_embedding[0].weight.detach()).abs().amax(dim=1) > 0
    assert torch.equal(moved, seen)
    with pytest.raises(IndexError):
        t.train(torch.zeros(B, G), torch.full((B,), 10), torch.zeros(B, dtype=torch.long))


def test_prefetch_gives_identical_training(tmp_path):
    """train(batch, prefetch=next batch) must be a pure scheduling change."""
    batches = [synthetic_tensors("paper", B, G, 5, 3, text_dim=24, patch_dim=40, seed=s, ragged=True) for s in (1, 2, 3)]
    host = [tuple(x.pin_memory() for x in (b[2], b[0], b[1], b[3], b[4])) for b in batches]   # train() argument order
    curves = []
    for use_prefetch in (False, True):
        torch.manual_seed(5)
        t, _ = _make("paper", tmp_path)
        t.build_WGAN_GP()
        t.init_train()
        t.dropout_p = 0.0
        torch.manual_seed(9)
        losses = []
        for i, b in enumerate(host):
            nxt = host[i + 1] if (use_prefetch and i + 1 < len(host)) else None
            t.train(*b, prefetch=nxt)
            losses.append((t.d_batch_loss.copy(), t.g_batch_loss.copy()))
        curves.append(losses)
    for (d0, g0), (d1, g1) in zip(*curves):
        np.testing.assert_array_equal(d0, d1)
        np.testing.assert_array_equal(g0, g1)
