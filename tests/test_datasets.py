"""On-disk dataset readers (gemmgan_b200/datasets.py + the two drop-in loader modules) on a small synthetic
dataset directory: tuple layouts / dtypes / masks, splits, normalisation, label encoding, and — when the reference
tree is present (build container) — tensor-for-tensor equality with the reference's own loaders under the same seed."""
import os
import pickle
import sys

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import ref_shim

N_CASES, N_GENES, DT, DP, T = 30, 12, 6, 5, 4


@pytest.fixture()
def dataset_dir(tmp_path, monkeypatch):
    rng = np.random.default_rng(0)
    cases = [f"case{i:02d}" for i in range(N_CASES)]
    expr = rng.gamma(2.0, 1.0, size=(N_CASES, N_GENES)).astype(np.float64)
    expr[:, 3] = 0.0                        # an all-zero gene: removed by the > 90 % zeros filter
    expr[rng.random((N_CASES, N_GENES)) < 0.2] = 0.0
    expr[:, 3] = 0.0
    pd.DataFrame(expr, index=cases, columns=[f"g{j}" for j in range(N_GENES)]).to_parquet(tmp_path / "rna_seq.parquet")
    (tmp_path / "case_ids.txt").write_text("\n".join(cases[:-1]) + "\n")   # the last case is not listed
    pd.DataFrame(rng.normal(size=(N_CASES, DT)), index=cases).to_parquet(tmp_path / "text.parquet")
    (tmp_path / "patches").mkdir()
    (tmp_path / "tokens").mkdir()
    for i, c in enumerate(cases):
        np.save(tmp_path / "patches" / f"{c}.npy", rng.normal(size=(1 + i % 7, DP)))   # 1..7 patches per case
        np.save(tmp_path / "tokens" / f"{c}.npy", rng.normal(size=(1, T, DT)))
        att = np.zeros((1, T), dtype=np.int64)
        att[0, :1 + i % T] = 1
        np.save(tmp_path / "tokens" / f"{c}_attention_mask.npy", att)
    meta = {c: dict(disease_type=f"d{i % 3}", primary_site=f"s{i % 4}") for i, c in enumerate(cases)}
    with open(tmp_path / "metainfos.pkl", "wb") as f:
        pickle.dump(meta, f)
    monkeypatch.chdir(tmp_path)   # the multi-patch loader writes gene_names.npy into the working directory
    return tmp_path


KW = dict(num_patches=4, batch_size=5, num_workers=0, text_embedding_file="text.parquet", patch_embeddings_folder="patches")


def test_multi_patch_tuples(dataset_dir):
    import multi_patch_gan_dataloader as m

    train, val, test, n_genes = m.dataloader_multi_patch_conditional_gan(dataset_dir, **KW)
    assert n_genes == N_GENES - 1
    n = N_CASES - 1
    assert (len(train.dataset), len(val.dataset), len(test.dataset)) == (int(0.64 * n), int(0.16 * n), n - int(0.64 * n) - int(0.16 * n))
    text, genes, patches, pad, disease, site = next(iter(test))
    assert text.shape == (5, DT) and genes.shape == (5, n_genes) and patches.shape == (5, 4, DP) and pad.shape == (5, 4)
    assert text.dtype == genes.dtype == patches.dtype == torch.float32 and pad.dtype == torch.bool
    assert disease.dtype == site.dtype == torch.long and int(disease.max()) <= 2 and int(site.max()) <= 3
    assert not pad.any()      # reference quirk (datasets.MASK_ZERO_PADDING): zero rows are NOT marked as padding
    short = (patches.abs().sum(-1) == 0)
    assert short.any() and not short[:, 0].any()                   # cases with < 4 patches got zero rows behind them
    # training genes are z-scored with the training statistics
    g = np.stack([train.dataset[i][1].numpy() for i in range(len(train.dataset))])
    assert np.allclose(g.mean(0), 0, atol=1e-5) and np.allclose(g.std(0)[g.std(0) > 0], 1, atol=1e-4)
    assert os.path.exists("gene_names.npy")


def test_mask_zero_padding_switch(dataset_dir, monkeypatch):
    from gemmgan_b200 import datasets

    monkeypatch.setattr(datasets, "MASK_ZERO_PADDING", True)
    _, _, test, _ = datasets.multi_patch_loaders(dataset_dir, **KW)
    _, _, patches, pad, _, _ = next(iter(test))
    assert pad.any() and (patches[pad] == 0).all() and not pad[:, 0].any()
    assert torch.equal(pad, patches.abs().sum(-1) == 0)


def test_multi_patch_multi_token_tuples(dataset_dir):
    import multi_patch_multi_token_gan_dataloader as m

    train, val, test, n_genes = m.dataloader_multi_patch_conditional_gan(dataset_dir, token_embeddings_folder="tokens", **KW)
    tokens, tpad, genes, patches, pad, disease, site = next(iter(train))
    assert tokens.shape == (5, T, DT) and tpad.shape == (5, T) and tpad.dtype == torch.bool
    assert genes.shape == (5, n_genes) and patches.shape == (5, 4, DP) and pad.shape == (5, 4)
    assert not tpad[:, 0].any() and disease.dtype == torch.long     # token 0 always attended; True = padding


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("which", ["multi_patch_gan_dataloader", "multi_patch_multi_token_gan_dataloader"])
def test_matches_reference_loader(dataset_dir, which):
    """Same seeds, same directory: every batch of every split equals the reference loader's (num_workers=0)."""
    kw = dict(KW)
    if "token" in which:
        kw["token_embeddings_folder"] = "tokens"
    ours_mod = __import__(which)
    ours = ours_mod.dataloader_multi_patch_conditional_gan(dataset_dir, **kw)
    ours_batches = [[b for b in loader] for loader in ours[:3]]
    sys.modules.pop(which, None)
    ref_mod = ref_shim.load(which)           # /root/reference/src/<which>.py
    assert ref_mod.__file__ != ours_mod.__file__
    ref = ref_mod.dataloader_multi_patch_conditional_gan(dataset_dir, **kw)
    assert ours[3] == ref[3]
    for ob, loader in zip(ours_batches, ref[:3]):
        rb = [b for b in loader]
        assert len(ob) == len(rb)
        for x, y in zip(ob, rb):
            for a, b in zip(x, y):
                assert a.dtype == b.dtype and torch.equal(a, b)
    sys.modules.pop(which, None)


@pytest.fixture()
def split_tables(dataset_dir):
    """The two tables the gene-only / label loaders intersect the cases with (hard-coded names, src/data_loader.py:109-110)."""
    import shutil

    shutil.copy(dataset_dir / "text.parquet", dataset_dir / "text_embeddings_contrastive_256.parquet")
    shutil.copytree(dataset_dir / "patches", dataset_dir / "patch_embeddings_contrastive_256")
    return dataset_dir


def test_gene_only_and_label_loader_tuples(split_tables):
    import benchmark_gan_dataloader as b
    import data_loader as d

    train, val, test, n_genes = d.dataloader_tcga(dataset_path=split_tables, batch_size=5, num_workers=0, seed=42)
    n = N_CASES - 1
    assert n_genes == N_GENES - 1 and len(test.dataset) == n - int(0.64 * n) - int(0.16 * n)
    (x,) = next(iter(test))
    assert x.shape == (5, n_genes) and x.dtype == torch.float64      # the trainer casts (vanilla_gan_unconditional.py:424)
    train, val, test, n_genes = b.dataloader_benchmark_conditional_gan(dataset_path=split_tables, batch_size=5,
                                                                     num_workers=0, seed=42)
    genes, disease, site = next(iter(train))
    assert genes.shape == (5, n_genes) and genes.dtype == torch.float32
    assert disease.dtype == site.dtype == torch.long and int(disease.max()) <= 2 and int(site.max()) <= 3
    assert isinstance(test.dataset, b.BenchmarkGANDataset) and len(val.dataset) == int(0.16 * n)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("which,fn", [("data_loader", "dataloader_tcga"),
                                      ("benchmark_gan_dataloader", "dataloader_benchmark_conditional_gan")])
def test_gene_only_and_label_loaders_match_the_reference(split_tables, which, fn):
    kw = dict(dataset_path=split_tables, batch_size=5, num_workers=0, seed=7)
    for name in ("data_loader", "benchmark_gan_dataloader"):
        sys.modules.pop(name, None)
    ours_mod = __import__(which)
    ours = getattr(ours_mod, fn)(**kw)
    ours_batches = [[b for b in loader] for loader in ours[:3]]
    for name in ("data_loader", "benchmark_gan_dataloader"):
        sys.modules.pop(name, None)
    ref_mod = ref_shim.load(which)
    assert ref_mod.__file__ != ours_mod.__file__
    ref = getattr(ref_mod, fn)(**kw)
    assert ours[3] == ref[3]
    for ob, loader in zip(ours_batches, ref[:3]):
        rb = [b for b in loader]
        assert len(ob) == len(rb)
        for x, y in zip(ob, rb):
            for a, b in zip(x, y):
                assert a.dtype == b.dtype and torch.equal(a, b)
    for name in ("data_loader", "benchmark_gan_dataloader"):
        sys.modules.pop(name, None)
