"""Device-side batch assembly on a B200: gemmgan_b200.datasets.DeviceResidentLoader (dataset resident
in HBM, gg_gather_rows per tensor) yields, tensor for tensor, the batches of the reference-layout Dataset read through
torch's DataLoader under the same numpy seed — patch sub-sampling (np.random.choice without replacement), zero
padding, masks, labels (src/multi_patch_multi_token_gan_dataloader.py:25-55, src/multi_patch_gan_dataloader.py:23-48)
— and a trainer consumes them."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from test_datasets import DP, DT, KW, T, dataset_dir  # noqa: F401  (fixture + the synthetic directory's sizes)
from gemmgan_b200 import datasets as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("module,num_patches", [("multi_patch_multi_token_gan_dataloader", 4),
                                                ("multi_patch_multi_token_gan_dataloader", 2),
                                                ("multi_patch_gan_dataloader", 3)])
def test_device_loader_equals_the_dataloader(dataset_dir, module, num_patches):
    m = __import__(module)
    kw = dict(KW, num_patches=num_patches)
    if "multi_token" in module:
        kw["token_embeddings_folder"] = "tokens"
    train, _, _, n_genes = m.dataloader_multi_patch_conditional_gan(dataset_dir, **kw)
    ds = train.dataset
    np.random.seed(123)
    want = list(DataLoader(ds, batch_size=5, shuffle=False, num_workers=0))
    np.random.seed(123)
    got = list(D.DeviceResidentLoader(ds, batch_size=5))
    assert len(got) == len(want) and len(want[-1][0]) == len(ds) % 5      # the last partial batch is kept
    for bw, bg in zip(want, got):
        assert len(bw) == len(bg)
        for tw, tg in zip(bw, bg):
            assert tg.is_cuda and tg.dtype == tw.dtype and tg.shape == tw.shape
            assert torch.equal(tg.cpu(), tw)
    # cases with more patches than num_patches were sub-sampled: a different seed picks other rows
    np.random.seed(7)
    other = list(D.DeviceResidentLoader(ds, batch_size=5))
    pi = 3 if "multi_token" in module else 2          # position of the patch tensor in the tuple
    assert any(not torch.equal(a[pi], b[pi]) for a, b in zip(got, other)) or num_patches >= 7


def _wide_dataset_dir(root, dt=8, dp=16, n_cases=30, n_genes=24, n_tokens=4):
    """Same directory layout as test_datasets.dataset_dir with feature widths the engine accepts (multiples of 8)."""
    import pickle

    import pandas as pd

    rng = np.random.default_rng(1)
    cases = [f"case{i:02d}" for i in range(n_cases)]
    expr = rng.gamma(2.0, 1.0, size=(n_cases, n_genes))
    pd.DataFrame(expr, index=cases, columns=[f"g{j}" for j in range(n_genes)]).to_parquet(root / "rna_seq.parquet")
    (root / "case_ids.txt").write_text("\n".join(cases) + "\n")
    pd.DataFrame(rng.normal(size=(n_cases, dt)), index=cases).to_parquet(root / "text.parquet")
    (root / "patches").mkdir()
    (root / "tokens").mkdir()
    for i, c in enumerate(cases):
        np.save(root / "patches" / f"{c}.npy", rng.normal(size=(1 + i % 7, dp)))
        np.save(root / "tokens" / f"{c}.npy", rng.normal(size=(1, n_tokens, dt)))
        att = np.zeros((1, n_tokens), dtype=np.int64)
        att[0, :1 + i % n_tokens] = 1
        np.save(root / "tokens" / f"{c}_attention_mask.npy", att)
    with open(root / "metainfos.pkl", "wb") as f:
        pickle.dump({c: dict(disease_type=f"d{i % 3}", primary_site=f"s{i % 4}") for i, c in enumerate(cases)}, f)
    return root


def test_trainer_consumes_device_batches(tmp_path, monkeypatch):
    import conditional_gan_cross_attention_with_film as paper
    import multi_patch_multi_token_gan_dataloader as m

    DT, DP = 8, 16
    data = tmp_path / "data"
    data.mkdir()
    _wide_dataset_dir(data, DT, DP)
    monkeypatch.chdir(tmp_path)
    kw = dict(KW, token_embeddings_folder="tokens")
    train, _, _, n_genes = m.dataloader_multi_patch_conditional_gan(data, **kw)
    loader = D.DeviceResidentLoader(train.dataset, batch_size=6, shuffle=True, generator=torch.Generator().manual_seed(0))
    torch.manual_seed(0)
    t = paper.WGAN_GP(input_dims=n_genes, optimizer="adam", results_dire=str(tmp_path), latent_dims=32, embedding_dims=32,
                      generator_dims=[32, 32, n_genes], discriminator_dims=[32, 32, 1], text_embedding_dims=DT,
                      patches_embedding_dims=DP)
    t.fit(loader, None, None, epochs=2)
    assert len(t.loss_dict["d loss"]) == 2 and np.isfinite(t.loss_dict["d loss"]).all()
