"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64 unless stated) of the reference's evaluation metrics
(SURVEY.md §8 row f4). Only tests/ may import this file; the product path (gemmgan_b200/evalmetrics.py) never does.

Pinned to the UNMODIFIED reference by oracle/make_eval_golden.py (fixtures tests/golden/eval_*.npz, checked by
tests/test_eval_oracle.py, and live against /root/reference when it is present). Every function names the reference
lines it follows. The third-party arithmetic behind the reference's calls is sklearn's
`pairwise_distances(metric='l1')` (= scipy cityblock, sum |a-b| in float64) and numpy partition / argpartition.
"""
from __future__ import annotations

import numpy as np


# ---- src/distribution_distances.py ---------------------------------------------------------------------------
def compute_pairwise_distance(data_x, data_y=None):
    """:51-66 — L1 distances [N, M] (sklearn pairwise_distances(metric='l1'))."""
    x = np.asarray(data_x, dtype=np.float64)
    y = x if data_y is None else np.asarray(data_y, dtype=np.float64)
    out = np.empty((x.shape[0], y.shape[0]), dtype=np.float64)
    for i in range(x.shape[0]):  # row blocks keep the broadcast small
        out[i] = np.abs(x[i][None, :] - y).sum(axis=1)
    return out


def get_kth_value(unsorted, k, axis=-1):
    """:69-83 — max of the k smallest = sorted[k-1]."""
    return np.sort(np.asarray(unsorted), axis=axis).take(k - 1, axis=axis)


def compute_nearest_neighbour_distances(input_features, nearest_k):
    """:86-99 — k+1 because the row holds the zero self-distance."""
    return get_kth_value(compute_pairwise_distance(input_features), k=nearest_k + 1, axis=-1)


def compute_prdc(real_features, fake_features, nearest_k):
    """:102-142."""
    r_real = compute_nearest_neighbour_distances(real_features, nearest_k)
    r_fake = compute_nearest_neighbour_distances(fake_features, nearest_k)
    d = compute_pairwise_distance(real_features, fake_features)
    inside_real = d < r_real[:, None]
    precision = inside_real.any(axis=0).mean()
    recall = (d < r_fake[None, :]).any(axis=1).mean()
    density = (1.0 / float(nearest_k)) * inside_real.sum(axis=0).mean()
    coverage = (d.min(axis=1) < r_real).mean()
    return dict(precision=precision, recall=recall, density=density, coverage=coverage)


# ---- src/unsupervised_metrics.py -----------------------------------------------------------------------------
def batch_pairwise_distances(U, V):
    """:114-138 — squared Euclidean distances, clamped at 0 (the reference expands |u|^2 - 2uv + |v|^2 in fp32)."""
    u = np.asarray(U, dtype=np.float64)
    v = np.asarray(V, dtype=np.float64)
    out = np.empty((u.shape[0], v.shape[0]), dtype=np.float64)
    for i in range(u.shape[0]):
        t = u[i][None, :] - v
        out[i] = (t * t).sum(axis=1)
    return out


class ManifoldEstimator:
    """:141-245 — radii D[:, q] = distance to the nhood_sizes[q]-th neighbour (rank counted with the zero
    self-distance at rank 0), membership test `distance <= D`."""

    def __init__(self, features, nhood_sizes=(3,), clamp_to_percentile=None, eps=1e-5):
        self.nhood_sizes = list(nhood_sizes)
        self.eps = eps
        self._ref = np.asarray(features, dtype=np.float64)
        d = np.sort(batch_pairwise_distances(self._ref, self._ref), axis=1)
        self.D = d[:, self.nhood_sizes].astype(np.float32)
        if clamp_to_percentile is not None:
            mx = np.percentile(self.D, clamp_to_percentile, axis=0)
            self.D[self.D > mx] = 0

    def evaluate(self, eval_features, return_realism=False, return_neighbors=False):
        d = batch_pairwise_distances(eval_features, self._ref).astype(np.float32)
        pred = np.any(d[:, :, None] <= self.D, axis=1).astype(np.int32)
        realism = np.max(self.D[:, 0] / (d + self.eps), axis=1)
        nearest = np.argmin(d, axis=1).astype(np.int32)
        if return_realism and return_neighbors:
            return pred, realism, nearest
        if return_realism:
            return pred, realism
        if return_neighbors:
            return pred, nearest
        return pred


def get_precision_recall(real_data, fake_data, nb_nn=(10,)):
    """:247-324 (knn_precision_recall_features + get_precision_recall)."""
    ref_m = ManifoldEstimator(real_data, nb_nn)
    eval_m = ManifoldEstimator(fake_data, nb_nn)
    precision = ref_m.evaluate(fake_data).mean(axis=0)
    recall = eval_m.evaluate(real_data).mean(axis=0)
    return precision[0], recall[0]


# ---- src/privacy_evaluator.py --------------------------------------------------------------------------------
def _l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.sqrt(batch_pairwise_distances(a, b))


def dcr(real_data, gen_data, test_data):
    """:9-32 — share of generated rows closer to the training set than to the test set."""
    d_real = _l2(gen_data, real_data).min(axis=1)
    d_test = _l2(gen_data, test_data).min(axis=1)
    return float((d_real < d_test).sum()) / d_real.shape[0]


def nndr_ratios(gen_data, other):
    s = np.sort(_l2(gen_data, other), axis=1)
    return s[:, 0] / s[:, 1]


def nndr(real_data, gen_data, test_data):
    """:34-66 — first / second neighbour distance ratio, train vs test."""
    a = nndr_ratios(gen_data, real_data)
    b = nndr_ratios(gen_data, test_data)
    return float((a < b).sum()) / a.shape[0]


# ---- src/corr_score.py ---------------------------------------------------------------------------------------
def upper_diag_list(m_):
    """:20-40 — strict upper triangle, row by row."""
    m = np.asarray(m_)
    iu = np.triu_indices(m.shape[0], k=1)
    return m[iu]


def standardize(a):
    """:55-61 — constant columns: 0/0 -> NaN -> replaced by a - mean (= 0)."""
    a = np.asarray(a, dtype=np.float64)
    off = a.mean(axis=0)
    sd = a.std(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (a - off) / sd
    bad = np.isnan(s)
    s[bad] = (a - off)[bad]
    return s


def pearson_correlation(x, y):
    """:43-68."""
    x = np.asarray(x)
    y = np.asarray(y)
    assert x.shape[0] == y.shape[0]
    return np.dot(standardize(x).T, standardize(y)) / x.shape[0]


def correlations_list(x, y):
    """:91-104."""
    return upper_diag_list(pearson_correlation(x, y))


def gamma_coef(x, y):
    """:106-120 (= gamma_coeff_score :71-88)."""
    dx = 1 - correlations_list(x, x)
    dy = 1 - correlations_list(y, y)
    return pearson_correlation(dx, dy)


def gamma_from_moments(m):
    """What the product's host side does with gg_gamma_moments' six sums (count, Sa, Sb, Saa, Sbb, Sab)."""
    n, sa, sb, saa, sbb, sab = (float(v) for v in m)
    ma, mb = sa / n, sb / n
    va, vb = saa / n - ma * ma, sbb / n - mb * mb
    return (sab / n - ma * mb) / np.sqrt(va * vb)
