"""Pins oracle/evalmetrics_ref.py (the CPU restatement of the reference's evaluation metrics) to the
committed fixtures tests/golden/eval_*.npz that oracle/make_eval_golden.py produced by running the UNMODIFIED
reference, and — in the build container, where /root/reference exists — live against the reference functions."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import evalmetrics_ref as ref
from oracle import ref_shim


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.mark.parametrize("name", ["eval_prdc_a", "eval_prdc_b"])
def test_prdc_matches_reference(name):
    fx = load(name)
    k = int(fx["k"])
    np.testing.assert_allclose(ref.compute_pairwise_distance(fx["real"], fx["fake"]), fx["dist"], rtol=1e-6)  # sklearn keeps fp32 inputs in fp32
    np.testing.assert_allclose(ref.compute_nearest_neighbour_distances(fx["real"], k), fx["radii_real"], rtol=1e-6)
    np.testing.assert_allclose(ref.compute_nearest_neighbour_distances(fx["fake"], k), fx["radii_fake"], rtol=1e-6)
    got = ref.compute_prdc(fx["real"], fx["fake"], k)
    for key in ("precision", "recall", "density", "coverage"):
        assert got[key] == pytest.approx(float(fx[key]), abs=1e-12), key


def test_get_kth_value_counts_duplicates():
    a = np.array([[3.0, 1.0, 1.0, 2.0, 5.0], [0.0, 0.0, 0.0, 7.0, 7.0]])
    assert ref.get_kth_value(a, 1).tolist() == [1.0, 0.0]
    assert ref.get_kth_value(a, 2).tolist() == [1.0, 0.0]
    assert ref.get_kth_value(a, 3).tolist() == [2.0, 0.0]
    assert ref.get_kth_value(a, 4).tolist() == [3.0, 7.0]


def test_knn_precision_recall_matches_reference():
    fx = load("eval_knn_pr")
    k = int(fx["k"])
    # the reference expands |u|^2 - 2uv + |v|^2 in fp32; the restatement takes differences in fp64
    np.testing.assert_allclose(ref.batch_pairwise_distances(fx["fake"], fx["real"]), fx["sqdist"], rtol=2e-5, atol=2e-4)
    est = ref.ManifoldEstimator(fx["real"], nhood_sizes=[k])
    np.testing.assert_allclose(est.D, fx["radii"], rtol=2e-5, atol=2e-4)
    pred, realism, nearest = est.evaluate(fx["fake"], return_realism=True, return_neighbors=True)
    assert np.array_equal(pred, fx["pred"])
    assert np.array_equal(nearest, fx["nearest"])
    np.testing.assert_allclose(realism, fx["realism"], rtol=1e-4)
    p, r = ref.get_precision_recall(fx["real"], fx["fake"], nb_nn=[k])
    assert p == pytest.approx(float(fx["precision"])) and r == pytest.approx(float(fx["recall"]))
    clamped = ref.ManifoldEstimator(fx["real"], clamp_to_percentile=50).evaluate(fx["fake"], return_realism=True)[1]
    np.testing.assert_allclose(clamped, fx["realism_clamped"], rtol=1e-4)


def test_privacy_scores_match_reference():
    fx = load("eval_privacy")
    assert ref.dcr(fx["real"], fx["fake"], fx["test"]) == pytest.approx(float(fx["dcr"]), abs=1e-12)
    assert ref.nndr(fx["real"], fx["fake"], fx["test"]) == pytest.approx(float(fx["nndr"]), abs=1e-12)


def test_gamma_matches_reference():
    fx = load("eval_gamma")
    corr = ref.pearson_correlation(fx["x"], fx["y2"])
    np.testing.assert_allclose(corr, fx["corr"], atol=2e-6)       # the reference works in fp32
    assert np.all(corr[7] == 0.0)                                  # constant gene: NaN -> 0 (corr_score.py:59)
    np.testing.assert_allclose(ref.correlations_list(fx["x"], fx["x"]), fx["corr_list"], atol=2e-6)
    assert ref.gamma_coef(fx["x"], fx["y"]) == pytest.approx(float(fx["gamma"]), abs=2e-6)
    assert float(fx["gamma_score"]) == pytest.approx(float(fx["gamma"]), abs=1e-7)
    # the six-moment closed form the product's host side evaluates
    a = ref.correlations_list(fx["x"], fx["x"])
    b = ref.correlations_list(fx["y"], fx["y"])
    m = [a.size, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum()]
    assert ref.gamma_from_moments(m) == pytest.approx(ref.gamma_coef(fx["x"], fx["y"]), abs=1e-12)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_live_against_reference_functions():
    dd = ref_shim.load("distribution_distances")
    cs = ref_shim.load("corr_score")
    r = np.random.RandomState(11)
    real = r.randn(45, 19).astype(np.float32)
    fake = (r.randn(38, 19) * 1.1 + 0.2).astype(np.float32)
    want = dd.compute_prdc(real, fake, 4)
    got = ref.compute_prdc(real, fake, 4)
    for key in want:
        assert got[key] == pytest.approx(want[key], abs=1e-12), key
    x = r.randn(25, 30).astype(np.float32)
    y = (x[:20] + 0.7 * r.randn(20, 30)).astype(np.float32)
    assert ref.gamma_coef(x, y) == pytest.approx(float(cs.gamma_coef(x, y)), abs=2e-6)


def test_product_module_has_the_reference_names_and_no_fallback():
    """CPU part of the drop-in check: every reference function name exists, and without a CUDA device the product
    raises instead of computing on the host."""
    import torch

    from gemmgan_b200 import _lib, evalmetrics as em

    for name in ("compute_pairwise_distance", "get_kth_value", "compute_nearest_neighbour_distances", "compute_prdc",
                 "batch_pairwise_distances", "ManifoldEstimator", "knn_precision_recall_features",
                 "get_precision_recall", "get_realism_score", "dcr", "nndr", "upper_diag_list", "pearson_correlation",
                 "correlations_list", "gamma_coef", "gamma_coeff_score"):
        assert callable(getattr(em, name)), name
    imports = [ln for ln in open(em.__file__).read().splitlines() if ln.lstrip().startswith(("import ", "from "))]
    assert not any("oracle" in ln or "sklearn" in ln or "scipy" in ln for ln in imports), imports
    if not torch.cuda.is_available():
        with pytest.raises(_lib.GGError):
            em.compute_prdc(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), 2)
