"""Checks of the evaluation-metric kernels (gemmgan_b200/csrc/evalmetrics.cu) through the host mirror
gemmgan_b200/evalmetrics.py, shared by two test modules that provide the `host` fixture:

  tests/test_eval_kernels_emulated.py  the kernels compiled for the host (tests/cuda_emu), CPU suite
  tests/test_gpu_zeval.py              the real library on a B200 (-m gpu)

Expected values come from the oracle (oracle/evalmetrics_ref.py) and from the fixtures tests/golden/eval_*.npz that
the unmodified reference produced. Tolerances: fp32 accumulation over <= 203 features against fp64 -> rtol 2e-6 .. 5e-6;
rank selection, membership flags, hit counts and the DCR / NNDR shares are exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gemmgan_b200 import _lib
from gemmgan_b200 import evalmetrics as em
from oracle import evalmetrics_ref as ref


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


DEV = {"device": "cpu"}   # set by the importing test module's `host` fixture


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV["device"])


def cpu(x):
    return x.cpu().numpy()


# ------------------------------------------------------------------------------------------ single kernels
@pytest.mark.parametrize("metric", [em.DIST_L1, em.DIST_SQL2, em.DIST_L2])
@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (5, 3, 7), (64, 64, 16), (65, 70, 33), (130, 3, 203)])
def test_pairwise_distance_kernel(host, metric, n, m, d):
    r = np.random.RandomState(n * 1000 + m * 10 + d)
    x, y = r.randn(n, d).astype(np.float32), r.randn(m, d).astype(np.float32)
    got = host.pairwise_distance(t(x), t(y), metric).cpu().numpy()
    diff = x.astype(np.float64)[:, None, :] - y.astype(np.float64)[None, :, :]
    want = np.abs(diff).sum(2) if metric == em.DIST_L1 else (diff * diff).sum(2)
    if metric == em.DIST_L2:
        want = np.sqrt(want)
    np.testing.assert_allclose(got, want, rtol=5e-6, atol=1e-6)


def test_pairwise_distance_respects_pitches(host):
    r = np.random.RandomState(0)
    xb, yb = t(r.randn(9, 40)), t(r.randn(11, 24))
    x, y = xb[:, 3:20], yb[:, 5:22]                      # row pitch 40 / 24, 17 features, unaligned starts
    out = torch.full((9, 16), -1.0, device=DEV["device"])
    L = _lib.lib()
    _lib.check(L.gg_pairwise_distance(C.c_void_p(x.data_ptr()), 40, C.c_void_p(y.data_ptr()), 24, 9, 11, 17,
                                      em.DIST_L1, C.c_void_p(out.data_ptr()), 16, None))
    np.testing.assert_allclose(out[:, :11].cpu().numpy(), ref.compute_pairwise_distance(x.cpu().numpy(), y.cpu().numpy()), rtol=2e-6)
    assert torch.all(out[:, 11:] == -1.0)               # nothing written past m


@pytest.mark.parametrize("n,m", [(3, 1), (4, 5), (7, 300), (2, 1000)])
def test_row_kth_smallest_kernel(host, n, m):
    r = np.random.RandomState(m)
    a = np.round(r.randn(n, m) * 3).astype(np.float32)   # rounded: many duplicates
    a[0, : min(m, 3)] = -np.inf if m > 3 else a[0, : min(m, 3)]
    s = np.sort(a, axis=1)
    for k in sorted({0, 1, m // 2, m - 1} & set(range(m))):
        kth, arg = host.row_kth_smallest(t(a), k, want_argmin=True)
        assert np.array_equal(kth.cpu().numpy(), s[:, k]), k
        assert np.array_equal(arg.cpu().numpy(), np.argmin(a, axis=1)), k
    with pytest.raises(_lib.GGError):
        host.row_kth_smallest(t(a), m)                   # rank outside the row


def test_row_membership_and_col_hits_kernels(host):
    r = np.random.RandomState(3)
    n, m = 37, 301
    d = np.abs(r.randn(n, m)).astype(np.float32)
    d[5, 17] = d[5, 200] = d[5].min() / 2                # tied minimum: first index wins
    col_r = np.abs(r.randn(m)).astype(np.float32) * 0.05
    row_r = np.abs(r.randn(n)).astype(np.float32) * 0.3
    col_r[9] = d[2, 9]                                   # exact tie: < excludes it, <= includes it
    for inclusive in (False, True):
        res = host.row_membership(t(d), t(col_r[None])[0], inclusive, eps=1e-5)
        cmp = (d <= col_r[None]) if inclusive else (d < col_r[None])
        assert np.array_equal(res["any"].cpu().numpy().astype(bool), cmp.any(1))
        assert np.array_equal(res["min"].cpu().numpy(), d.min(1))
        assert np.array_equal(res["argmin"].cpu().numpy(), d.argmin(1))
        np.testing.assert_allclose(res["ratio"].cpu().numpy(), (col_r[None] / (d + np.float32(1e-5))).max(1), rtol=1e-6)
        hits = torch.zeros(m, dtype=torch.int32, device=DEV["device"])
        host.col_hits(t(d), t(row_r[None])[0], inclusive, hits)
        host.col_hits(t(d), t(row_r[None])[0], inclusive, hits)          # accumulates
        cmp = (d <= row_r[:, None]) if inclusive else (d < row_r[:, None])
        assert np.array_equal(hits.cpu().numpy(), 2 * cmp.sum(0))
    only_min = host.row_membership(t(d), None, False, want=("min", "argmin"))
    assert only_min["any"] is None and np.array_equal(only_min["min"].cpu().numpy(), d.min(1))


def test_standardize_and_correlation_kernels(host):
    fx = load("eval_gamma")
    xs = host.standardize_columns(t(fx["x"])).cpu().numpy()
    np.testing.assert_allclose(xs, ref.standardize(fx["x"]), atol=2e-6)
    assert np.all(xs[:, 7] == 0.0)                       # constant gene
    corr = host.pearson_correlation(fx["x"], fx["y2"])
    np.testing.assert_allclose(corr, fx["corr"], atol=3e-6)
    np.testing.assert_allclose(host.correlations_list(fx["x"], fx["x"]), fx["corr_list"], atol=3e-6)
    a, b = fx["x"][:, 3], fx["y2"][:, 11]
    assert host.pearson_correlation(a, b) == pytest.approx(np.corrcoef(a, b)[0, 1], abs=2e-6)   # 1-D lists


@pytest.mark.parametrize("g", [2, 50, 64, 65, 130])
def test_gamma_moments_kernel(host, g):
    r = np.random.RandomState(g)
    mix = r.randn(6, g)
    x = (r.randn(21, 6) @ mix + 0.5 * r.randn(21, g)).astype(np.float32)
    y = (r.randn(17, 6) @ mix + 0.8 * r.randn(17, g)).astype(np.float32)
    m = host.gamma_moments(x, y)
    a, b = ref.correlations_list(x, x), ref.correlations_list(y, y)
    want = [a.size, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum()]
    np.testing.assert_allclose(m, want, rtol=1e-5, atol=1e-5)
    if g > 2:
        assert host.gamma_coef(x, y) == pytest.approx(float(ref.gamma_coef(x, y)), abs=5e-6)


# ------------------------------------------------------------- reference function by reference function
@pytest.mark.parametrize("name", ["eval_prdc_a", "eval_prdc_b"])
def test_prdc_against_reference_golden(host, name):
    fx = load(name)
    k = int(fx["k"])
    np.testing.assert_allclose(host.compute_pairwise_distance(fx["real"], fx["fake"]), fx["dist"], rtol=2e-6)
    np.testing.assert_allclose(host.compute_nearest_neighbour_distances(fx["real"], k), fx["radii_real"], rtol=2e-6)
    np.testing.assert_allclose(host.get_kth_value(fx["dist"], k + 1), np.sort(fx["dist"], axis=1)[:, k])
    got = host.compute_prdc(fx["real"], fx["fake"], k)
    for key in ("precision", "recall", "density", "coverage"):
        assert got[key] == pytest.approx(float(fx[key]), abs=1e-9), key


def test_prdc_is_chunk_invariant(host, monkeypatch):
    fx = load("eval_prdc_b")
    whole = host.compute_prdc(fx["real"], fx["fake"], 10)
    monkeypatch.setattr(em, "CHUNK_BYTES", 64 * 4 * 97)  # 64 rows of the [130, 97] matrix at a time
    assert host.compute_prdc(fx["real"], fx["fake"], 10) == whole


def test_knn_precision_recall_against_reference_golden(host):
    fx = load("eval_knn_pr")
    k = int(fx["k"])
    np.testing.assert_allclose(host.batch_pairwise_distances(fx["fake"], fx["real"]).cpu().numpy(), fx["sqdist"],
                               rtol=2e-5, atol=2e-4)     # the reference's fp32 |u|^2 - 2uv + |v|^2 cancels
    est = host.ManifoldEstimator(fx["real"], nhood_sizes=[k])
    np.testing.assert_allclose(est.D, fx["radii"], rtol=2e-5, atol=2e-4)
    pred, realism, nearest = est.evaluate(fx["fake"], return_realism=True, return_neighbors=True)
    assert pred.dtype == np.int32 and pred.shape == fx["pred"].shape and np.array_equal(pred, fx["pred"])
    assert np.array_equal(nearest, fx["nearest"])
    np.testing.assert_allclose(realism, fx["realism"], rtol=1e-4)
    p, r = host.get_precision_recall(fx["real"], fx["fake"], nb_nn=[k])
    assert p == pytest.approx(float(fx["precision"])) and r == pytest.approx(float(fx["recall"]))
    np.testing.assert_allclose(host.get_realism_score(fx["real"], fx["fake"]), fx["realism_clamped"], rtol=1e-4)


def test_privacy_scores_against_reference_golden(host):
    fx = load("eval_privacy")
    assert host.dcr(fx["real"], fx["fake"], fx["test"]) == pytest.approx(float(fx["dcr"]), abs=1e-12)
    assert host.nndr(fx["real"], fx["fake"], fx["test"]) == pytest.approx(float(fx["nndr"]), abs=1e-12)


def test_gamma_against_reference_golden(host):
    fx = load("eval_gamma")
    assert host.gamma_coef(fx["x"], fx["y"]) == pytest.approx(float(fx["gamma"]), abs=5e-6)
    assert host.gamma_coeff_score(fx["x"], fx["y"]) == pytest.approx(float(fx["gamma_score"]), abs=5e-6)


def test_argument_errors_are_reported(host):
    L = _lib.lib()
    buf = torch.zeros(4, 4, device=DEV["device"])
    p = C.c_void_p(buf.data_ptr())
    assert L.gg_pairwise_distance(p, 2, p, 4, 4, 4, 4, 0, p, 4, None) == -1      # ldx < d
    assert b"leading dimension" in L.gg_last_error()
    assert L.gg_pairwise_distance(p, 4, p, 4, 4, 4, 4, 7, p, 4, None) == -1      # unknown metric
    assert L.gg_gamma_moments(p, 4, 4, p, 4, 4, 4, None, 0, p, None) == -4        # GG_ERR_WORKSPACE
    with pytest.raises(ValueError):
        host.pairwise_distance(t(np.zeros((2, 3))), t(np.zeros((2, 4))), 0)


def test_evaluate_generated_composition(host):
    """TrainerBase.evaluate_generated (what fit()'s evaluation block computes on the GPU) against the oracle."""
    from gemmgan_b200.trainer import TrainerBase

    fx = load("eval_privacy")
    real, gen, test = fx["real"], fx["fake"], fx["test"]
    test_gen = gen[:35]
    got = TrainerBase.evaluate_generated(None, real, gen, test, test_gen, nn=4)
    train_want, test_want = ref.compute_prdc(real, gen, 4), ref.compute_prdc(test, test_gen, 4)
    for key in ("precision", "recall", "density", "coverage"):
        assert got[key] == pytest.approx(train_want[key], abs=1e-9), key
        assert got[key + "_test"] == pytest.approx(test_want[key], abs=1e-9), key
    assert got["gamma"] == pytest.approx(float(ref.gamma_coef(test, test_gen)), abs=5e-6)
    assert got["dcr"] == pytest.approx(float(fx["dcr"]), abs=1e-12) and got["nndr"] == pytest.approx(float(fx["nndr"]), abs=1e-12)


def test_generated_array_layout_and_privacy_report(host, tmp_path):
    """The twelve .npy files of the reference's test block and the DCR / NNDR report read back from them."""
    from gemmgan_b200 import trainer as tr

    fx = load("eval_privacy")
    real, gen, test = fx["real"], fx["fake"], fx["test"]
    lab = lambda a: np.arange(len(a)) % 7
    for run in range(2):
        tr.save_generated_arrays(os.path.join(tmp_path, f"test_{run}_epoch_5"),
                                 (real, gen[: len(real)], lab(real), lab(real), lab(real), lab(real)),
                                 (test, gen[:len(test)], lab(test), lab(test), lab(test), lab(test)))
    names = sorted(os.listdir(os.path.join(tmp_path, "test_0_epoch_5")))
    assert names == sorted(n + ".npy" for n in (
        "data_real", "data_gen", "test_real", "test_gen", "train_labels_real", "train_labels_gen", "test_labels_real",
        "test_labels_gen", "train_primary_site_real", "train_primary_site_gen", "test_primary_site_real",
        "test_primary_site_gen"))                                    # …with_film.py:795-806
    back = tr.load_generated_arrays(os.path.join(tmp_path, "test_1_epoch_5"))
    assert np.array_equal(back["data_real"], real) and np.array_equal(back["test_gen"], gen[:len(test)])
    rep = tr.privacy_report(str(tmp_path))
    want = ref.dcr(real, gen[: len(real)], test)
    assert rep["dcr"] == [want, want] and rep["std_dcr"] == 0.0
    assert rep["mean_nndr"] == pytest.approx(ref.nndr(real, gen[: len(real)], test), abs=1e-12)
    with pytest.raises(ValueError):
        tr.save_generated_arrays(str(tmp_path / "bad"), (real, gen), (test, gen))


def test_fit_evaluation_block(host, tmp_path):
    """TrainerBase._fit_evaluation (the GPU part of the reference fit()'s evaluation, …with_film.py:702-811) on a stub
    trainer whose generate_samples_all returns fixed arrays: per-epoch precision / recall / gamma at the test
    frequency, and the two saved runs with their metrics at the last epoch."""
    from gemmgan_b200.trainer import TrainerBase

    fx = load("eval_privacy")
    lab = lambda a: np.arange(len(a)) % 5
    six = lambda real, gen: (real, gen, lab(real), lab(real), lab(real), lab(real))
    loaders = {"train": six(fx["real"], fx["fake"][:70]), "val": six(fx["test"], fx["fake"][70:110]),
               "test": six(fx["test"], fx["fake"][110:150])}

    class Stub(TrainerBase):
        def __init__(self):
            self.result_dire, self.freq_compute_test = str(tmp_path), 2
            self.precision_scores, self.recall_scores, self.corr_scores = {}, {}, {}

        def generate_samples_all(self, loader):
            return loaders[loader]

    t = Stub()
    for epoch in range(4):
        t._fit_evaluation(epoch, 4, "train", "val", "test", val=True)
    assert sorted(t.precision_scores) == [2, 4] and sorted(t.corr_scores) == [2, 4]
    want = ref.compute_prdc(fx["test"], fx["fake"][70:110], 10)
    assert t.precision_scores[2] == pytest.approx(want["precision"], abs=1e-9)
    assert t.recall_scores[4] == pytest.approx(want["recall"], abs=1e-9)
    assert t.corr_scores[2] == pytest.approx(float(ref.gamma_coef(fx["test"], fx["fake"][70:110])), abs=5e-6)
    assert [os.path.basename(r["folder"]) for r in t.test_runs] == ["test_0_epoch_4", "test_1_epoch_4"]
    assert t.test_runs[0]["dcr"] == pytest.approx(ref.dcr(fx["real"], fx["fake"][:70], fx["test"]), abs=1e-12)
    assert np.array_equal(np.load(os.path.join(t.test_runs[1]["folder"], "test_gen.npy")), fx["fake"][110:150])
    quiet = Stub()
    quiet._fit_evaluation(3, 4, "train", None, None, val=True)        # no loaders: nothing happens
    quiet._fit_evaluation(3, 4, "train", "val", "test", val=False)
    assert quiet.precision_scores == {} and not hasattr(quiet, "test_runs")
