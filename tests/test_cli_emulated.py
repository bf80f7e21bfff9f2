"""`python conditional_gan_*.py --dataset_path ...` end to end on the CPU suite (gemmgan_b200/cli.py): the block every
reference script ends in (src/conditional_gan_cross_attention_with_film.py:902-995) — flags, the on-disk dataset through
the drop-in loader modules, WGAN_GP.fit(train, validation, test) with its periodic validation metrics, checkpoints and
the `test_<run>_epoch_<n>/` .npy dumps, then the DCR / NNDR privacy block — on the host-emulated engine and
evaluation kernels (tests/cuda_emu). The dataset directory has the reference's layout (rna_seq.parquet, case_ids.txt,
text table, per-case patch / token .npy files, metainfos.pkl; src/multi_patch_multi_token_gan_dataloader.py:58-187)."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

import emu_build
import host_trainer
from gemmgan_b200 import _abi_decl as A
from gemmgan_b200 import cli
from gemmgan_b200 import evalmetrics as em

N_CASES, N_GENES, DT, DP, T = 80, 40, 24, 32, 3      # validation split = 12 rows (the 10th neighbour needs 11)


class _Both:
    """Entry points of the engine build first, of the evaluation-kernel build otherwise (two emulated translation units)."""

    def __init__(self, *libs):
        self._libs = libs

    def __getattr__(self, name):
        for L in self._libs:
            try:
                return getattr(L, name)
            except AttributeError:
                continue
        raise AttributeError(name)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("cuda_emu")
    eng = emu_build.build("engine", d, cudart=True)
    A.declare(eng)
    ev = emu_build.build("evalmetrics", d)
    A.declare_evalmetrics(ev)
    return _Both(eng, ev)


@pytest.fixture()
def host(emu, monkeypatch):
    monkeypatch.setattr(em, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(em, "_stream", lambda: None)
    return host_trainer.apply(monkeypatch.setattr, emu)


@pytest.fixture()
def dataset_dir(tmp_path, monkeypatch):
    rng = np.random.default_rng(0)
    root = tmp_path / "data"
    root.mkdir()
    cases = [f"case{i:02d}" for i in range(N_CASES)]
    expr = rng.gamma(2.0, 1.0, size=(N_CASES, N_GENES)).astype(np.float64)
    pd.DataFrame(expr, index=cases, columns=[f"g{j}" for j in range(N_GENES)]).to_parquet(root / "rna_seq.parquet")
    (root / "case_ids.txt").write_text("\n".join(cases) + "\n")
    pd.DataFrame(rng.normal(size=(N_CASES, DT)), index=cases).to_parquet(root / "text.parquet")
    (root / "patches").mkdir()
    (root / "tokens").mkdir()
    for i, c in enumerate(cases):
        np.save(root / "patches" / f"{c}.npy", rng.normal(size=(1 + i % 7, DP)))   # 1..7 patches per case
        np.save(root / "tokens" / f"{c}.npy", rng.normal(size=(1, T, DT)))
        att = np.zeros((1, T), dtype=np.int64)
        att[0, :1 + i % T] = 1
        np.save(root / "tokens" / f"{c}_attention_mask.npy", att)
    meta = {c: dict(disease_type=f"d{i % 3}", primary_site=f"s{i % 4}") for i, c in enumerate(cases)}
    with open(root / "metainfos.pkl", "wb") as f:
        pickle.dump(meta, f)
    monkeypatch.chdir(tmp_path)   # the multi-patch loader writes gene_names.npy into the working directory
    return root


def flags(dataset_dir, out, extra=()):
    return ["--dataset_path", str(dataset_dir), "--output_path", str(out), "--num_epochs", "1", "--batch_size", "8",
            "--latent_dim", "16", "--hidden_dim", "32", "--embedding_dim", "32", "--num_patches", "4",
            "--num_workers", "0", "--freq_compute_test", "1", "--optimizer", "adam", "--text_embedding_file",
            "text.parquet", "--patch_embeddings_folder", "patches", "--token_embeddings_folder", "tokens",
            "--text_embedding_dims", str(DT), "--patches_embedding_dims", str(DP), *extra]


@pytest.mark.parametrize("script", ["paper", "film"])
def test_script_main_on_a_dataset_directory(host, script, dataset_dir, tmp_path, capsys):
    out = tmp_path / "run"
    model = cli.main(script, flags(dataset_dir, out))
    n_train, n_val = int(0.64 * N_CASES), int(0.16 * N_CASES)
    n_test = N_CASES - n_train - n_val
    # training: one epoch over ceil(51 / 8) batches; every loader ends in a partial batch with an engine of its own
    assert len(model.loss_dict["d loss"]) == 1 and np.isfinite(model.loss_dict["g loss"]).all()
    assert model.n_genes == N_GENES and sorted(model._engines) == sorted({8, n_train % 8, n_val % 8, n_test % 8} - {0})
    # validation metrics every freq_compute_test epochs (…with_film.py:702-734)
    assert sorted(model.precision_scores) == [1] and sorted(model.corr_scores) == [1]
    assert all(0.0 <= v <= 1.0 for v in model.precision_scores.values())
    # checkpoints (:710-711, :743-744) and the two final runs with their twelve arrays (:786-806)
    for name in ("generator_last_epoch.pt", "discriminator_last_epoch.pt"):
        assert (out / name).exists(), name
    for run in (0, 1):
        folder = out / f"test_{run}_epoch_1"
        assert np.load(folder / "data_real.npy").shape == (n_train, N_GENES)
        assert np.load(folder / "data_gen.npy").shape == (n_train, N_GENES)
        assert np.load(folder / "test_gen.npy").shape == (n_test, N_GENES)
        assert np.load(folder / "test_labels_real.npy").shape == (n_test,)
    # the two runs draw different noise
    assert not np.array_equal(np.load(out / "test_0_epoch_1" / "test_gen.npy"), np.load(out / "test_1_epoch_1" / "test_gen.npy"))
    # privacy block (:962-995) on the dumps, printed in the reference's format
    assert len(model.privacy["dcr"]) == 2 and 0.0 <= model.privacy["mean_nndr"] <= 1.0
    text = capsys.readouterr().out
    assert "Arguments: {" in text and "--------- Privacy Evaluation ----------" in text
    assert f"DCR {model.privacy['mean_dcr']:.4f}±{model.privacy['std_dcr']:.4f}, NNDR " in text
    assert "Best epoch correlation:" in text


def test_script_main_on_synthetic_batches(host, tmp_path):
    """No dataset directory: synthetic batches in the loader's tuple layout (what `python <script>.py` runs by default)."""
    model = cli.main("cross", ["--num_epochs", "1", "--batch_size", "4", "--latent_dim", "16", "--hidden_dim", "32",
                               "--embedding_dim", "32", "--num_patches", "5", "--num_text_tokens", "3", "--n_genes", "203",
                               "--text_embedding_dims", str(DT), "--patches_embedding_dims", str(DP),
                               "--synthetic_batches", "2", "--optimizer", "adam"])
    assert len(model.loss_dict["d loss"]) == 1 and np.isfinite(model.loss_dict["d loss"]).all()
    assert not hasattr(model, "privacy")
    import conditional_gan_concat as c
    assert c.parse_args(["--condition_type", "image"]).condition_type == "image"
    with pytest.raises(SystemExit):
        cli.build_parser("film").parse_args(["--condition_type", "image"])      # only the concat script has it


@pytest.mark.parametrize("script", ["vanilla", "label"])
def test_gene_only_and_label_scripts_on_a_dataset_directory(host, script, dataset_dir, tmp_path):
    """vanilla_gan_unconditional.py:777-799 / benchmark_generative_model.py:917-962: dataloader_tcga /
    dataloader_benchmark_conditional_gan, vocabulary sizes read from metainfos.pkl, fit(train, test, epochs)."""
    import shutil

    shutil.copy(dataset_dir / "text.parquet", dataset_dir / "text_embeddings_contrastive_256.parquet")
    shutil.copytree(dataset_dir / "patches", dataset_dir / "patch_embeddings_contrastive_256")
    out = tmp_path / "run"
    model = cli.main(script, ["--dataset_path", str(dataset_dir), "--output_path", str(out), "--batch_size", "16",
                              "--epochs", "2", "--latent_dim", "16", "--hidden_dim", "32", "--num_workers", "0"])
    assert len(model.loss_dict["d loss"]) == 2 and np.isfinite(model.loss_dict["d loss"]).all()
    assert model.n_genes == N_GENES and sorted(model._engines) == [int(0.64 * N_CASES) % 16, 16]
    if script == "label":
        assert model.vocab_sizes == [3, 4] and model.gen.categorical_embedded_dims == 256
    else:
        assert (out / "generator_last_epoch.pt").exists()          # (:614-615; the label script saves at its test epochs)


def test_device_resident_loader_feeds_the_script(host, dataset_dir, tmp_path):
    """--device_loader: every split uploaded once, batches assembled by gg_gather_rows — here in host
    memory by the emulated kernel. The batches equal the DataLoader's under the same numpy seed, and the script runs."""
    import multi_patch_multi_token_gan_dataloader as m
    from torch.utils.data import DataLoader

    from gemmgan_b200.datasets import DeviceResidentLoader

    train, _, _, _ = m.dataloader_multi_patch_conditional_gan(dataset_dir, num_patches=4, batch_size=8, num_workers=0,
                                                              text_embedding_file="text.parquet",
                                                              patch_embeddings_folder="patches",
                                                              token_embeddings_folder="tokens")
    np.random.seed(3)
    want = list(DataLoader(train.dataset, batch_size=8, shuffle=False, num_workers=0))
    np.random.seed(3)
    got = list(DeviceResidentLoader(train.dataset, batch_size=8))
    assert len(got) == len(want) == 7
    for bw, bg in zip(want, got):
        assert len(bw) == len(bg) == 7
        for tw, tg in zip(bw, bg):
            assert tg.dtype == tw.dtype and torch.equal(tg, tw)
    model = cli.main("paper", flags(dataset_dir, tmp_path / "run", ["--device_loader"]))
    assert len(model.loss_dict["d loss"]) == 1 and len(model.test_runs) == 2 and sorted(model.precision_scores) == [1]
