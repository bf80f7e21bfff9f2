"""Evaluation-metric kernels on the B200 through the C ABI: the shared cases of tests/eval_cases.py
(oracle + reference goldens) and size-independent properties at the reference's real shapes (18 868 genes).

Named to sort after the hot-path GPU tests: these kernels were written after the round's GPU minutes were spent and
had only been run through the host emulation (tests/test_eval_kernels_emulated.py) when they were committed."""
import numpy as np
import pytest
import torch

import eval_cases
from eval_cases import *  # noqa: F401,F403  (the shared test functions)
from gemmgan_b200 import _lib
from gemmgan_b200 import evalmetrics as em

pytestmark = pytest.mark.gpu


@pytest.fixture()
def host(monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _lib.require_device(torch.cuda.current_device())      # fails loudly off sm_100: there is no fallback
    monkeypatch.setitem(eval_cases.DEV, "device", "cuda")
    return em


def _profiles(n, g, seed, shift=0.0, scale=1.0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    centres = torch.randn(8, g, device="cuda", generator=gen) * 1.5
    pick = torch.randint(0, 8, (n,), device="cuda", generator=gen)
    return centres[pick] + scale * torch.randn(n, g, device="cuda", generator=gen) + shift


def test_distances_at_full_gene_count_against_torch_fp64(host):
    x, y = _profiles(300, 18868, 1), _profiles(257, 18868, 2, shift=0.1)
    for metric, p in ((em.DIST_L1, 1.0), (em.DIST_L2, 2.0)):
        got = host.pairwise_distance(x, y, metric)
        want = torch.cdist(x.double(), y.double(), p=p)   # independent check only (tests may use torch math)
        assert torch.allclose(got.double(), want, rtol=1e-4, atol=0)   # 18 868 sequential fp32 additions per pair
    sq = host.pairwise_distance(x, y, em.DIST_SQL2)
    assert torch.allclose(sq.double(), torch.cdist(x.double(), y.double()) ** 2, rtol=1e-4)
    self_d = host.pairwise_distance(x, x, em.DIST_L1)
    assert torch.all(torch.diagonal(self_d) == 0) and torch.equal(self_d, self_d.T.contiguous())


def test_rank_selection_matches_a_sort_at_scale(host):
    d = host.pairwise_distance(_profiles(513, 512, 3), _profiles(2049, 512, 4), em.DIST_L1)
    s, idx = torch.sort(d, dim=1, stable=True)
    for k in (0, 1, 10, 2048):
        kth, arg = host.row_kth_smallest(d, k, want_argmin=True)
        assert torch.equal(kth, s[:, k])
        assert torch.equal(arg.long(), idx[:, 0])


def test_prdc_properties_at_scale(host, monkeypatch):
    real = _profiles(1500, 18868, 5)
    same = host.compute_prdc(real, real, 10)               # a set against itself: everything is covered
    assert same["precision"] == 1.0 and same["recall"] == 1.0 and same["coverage"] == 1.0
    fake = _profiles(1100, 18868, 6, shift=0.05, scale=1.1)
    whole = host.compute_prdc(real, fake, 10)
    assert all(0.0 <= whole[k] <= 1.0 for k in ("precision", "recall", "coverage")) and whole["density"] >= 0.0
    monkeypatch.setattr(em, "CHUNK_BYTES", 256 * 4 * 1100)  # 256 distance rows at a time
    assert host.compute_prdc(real, fake, 10) == whole       # chunking does not change a single count
    far = host.compute_prdc(real, fake + 50.0, 10)           # disjoint supports
    assert far["precision"] == 0.0 and far["recall"] == 0.0 and far["coverage"] == 0.0 and far["density"] == 0.0


def test_privacy_scores_properties(host):
    real, test = _profiles(800, 18868, 7), _profiles(300, 18868, 8)
    copies = real[:200] + 1e-3 * _profiles(200, 18868, 9, scale=1.0)
    assert host.dcr(real, copies, test) == 1.0              # near copies of training rows are closer to train
    assert host.nndr(real, copies, test) == 1.0             # and their first neighbour is far closer than the second
    assert host.dcr(test, copies, real) == 0.0


def test_gamma_at_full_gene_count(host):
    x = _profiles(256, 18868, 10)
    assert host.gamma_coef(x, x) == pytest.approx(1.0, abs=1e-9)
    m = host.gamma_moments(x, x)
    g = 18868
    assert m[0] == g * (g - 1) // 2 and m[3] == pytest.approx(m[4]) and m[3] == pytest.approx(m[5])
    # against the unfused kernels: full [G, G] correlation of a gene subset, then the list statistics in fp64
    sub_x, sub_y = x[:, :700].contiguous(), _profiles(300, 700, 11)
    cx = torch.from_numpy(host.pearson_correlation(sub_x, sub_x)).double()
    cy = torch.from_numpy(host.pearson_correlation(sub_y, sub_y)).double()
    iu = torch.triu_indices(700, 700, offset=1)
    a, b = cx[iu[0], iu[1]], cy[iu[0], iu[1]]
    want = torch.corrcoef(torch.stack([a, b]))[0, 1].item()
    assert host.gamma_coef(sub_x, sub_y) == pytest.approx(want, abs=1e-6)
    # fp32 standardised values and a 256-term fp32 FMA dot per entry against fp64 corrcoef: |err| <= ~n * eps32 / 2
    # in the worst case (1.5e-5 at n = 256), a few 1e-6 in practice
    assert torch.allclose(cx, torch.corrcoef(sub_x.double().T).cpu(), atol=2e-5)


def test_evaluate_generated_after_fit(host, tmp_path):
    """fit() -> generate_samples_all (train and test loaders) -> TrainerBase.evaluate_generated: the arrays the
    reference hands to gamma_coef / compute_evaluation_metrics / dcr / nndr (…with_film.py:716-734, :983-984)."""
    import conditional_gan_cross_attention_with_film as paper
    from gemmgan_b200.synthetic import synthetic_loader

    G, B = 300, 16
    torch.manual_seed(0)
    t = paper.WGAN_GP(input_dims=G, optimizer="adam", results_dire=str(tmp_path), latent_dims=32, embedding_dims=32,
                      generator_dims=[32, 32, G], discriminator_dims=[32, 32, 1], text_embedding_dims=24,
                      patches_embedding_dims=40)
    mk = lambda n, seed: synthetic_loader("paper", n_samples=n, batch_size=B, n_genes=G, n_patches=5, n_tokens=3,
                                          seed=seed, text_dim=24, patch_dim=40, ragged=True)
    train, test = mk(4 * B, 1), mk(2 * B, 2)
    t.freq_compute_test = 1
    t.fit(train, test, test, epochs=1)                   # validation + test loaders: the evaluation block runs
    assert sorted(t.precision_scores) == [1] and sorted(t.corr_scores) == [1] and len(t.test_runs) == 2
    assert (tmp_path / "test_1_epoch_1" / "train_primary_site_gen.npy").exists()
    assert np.load(tmp_path / "test_0_epoch_1" / "test_real.npy").shape == (2 * B, G)
    assert 0.0 <= t.test_runs[0]["dcr"] <= 1.0 and -1.0 <= t.test_runs[0]["gamma"] <= 1.0
    data_real, data_gen = t.generate_samples_all(train)[:2]
    test_real, test_gen = t.generate_samples_all(test)[:2]
    got = t.evaluate_generated(data_real, data_gen, test_real, test_gen, nn=5)
    want = ref.compute_prdc(test_real, test_gen, 5)
    for key in ("precision", "recall", "density", "coverage"):
        assert got[key + "_test"] == pytest.approx(want[key], abs=1e-9), key
        assert 0.0 <= got[key]
    assert got["gamma"] == pytest.approx(float(ref.gamma_coef(test_real, test_gen)), abs=1e-5)
    assert got["dcr"] == pytest.approx(ref.dcr(data_real, data_gen, test_real), abs=1e-9)
    assert got["nndr"] == pytest.approx(ref.nndr(data_real, data_gen, test_real), abs=1e-9)
