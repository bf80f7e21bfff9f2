"""GPU evaluation metrics of the reference's fit() loop, behind the reference's own function
names, on the hand-written kernels of csrc/evalmetrics.cu (C ABI: include/gemmgan.h, "evaluation metrics").

Drop-in for (same names, argument meaning and return types; numpy in / numpy out unless the reference returns torch):

  src/distribution_distances.py:51-142   compute_pairwise_distance, get_kth_value,
                                         compute_nearest_neighbour_distances, compute_prdc
  src/unsupervised_metrics.py:114-345    batch_pairwise_distances, ManifoldEstimator, knn_precision_recall_features,
                                         get_precision_recall, get_realism_score
  src/privacy_evaluator.py:9-66          dcr, nndr
  src/corr_score.py:20-120               upper_diag_list, pearson_correlation, correlations_list, gamma_coef,
                                         gamma_coeff_score

PyTorch is device memory and streams only: every distance, neighbour rank, membership count, standardisation and
correlation is computed by the library's kernels; what remains on the host are means of [N]-sized flag vectors and the
closed-form gamma from six sums. There is no CPU fallback: without a CUDA device or the library these raise.

The [N, M] distance matrix is never held whole: rows are processed in chunks of at most `CHUNK_BYTES`, and
compute_prdc / gamma_coef never materialise what the reference does ([N, M] boolean cubes, two [G, G] matrices).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

DIST_L1, DIST_SQL2, DIST_L2 = 0, 1, 2
CHUNK_BYTES = 1 << 30  # distance rows resident at a time


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.GGError("gemmgan_b200.evalmetrics needs a CUDA device (sm_100); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    _lib.require_device(dev.index)
    return dev


def _f32(a, dev: torch.device) -> torch.Tensor:
    """numpy / torch, any float dtype -> contiguous fp32 2-D tensor on `dev`."""
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D [samples, features] array, got shape {tuple(t.shape)}")
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _row_chunks(n: int, m: int):
    rows = max(64, min(n, CHUNK_BYTES // (4 * max(m, 1)) // 64 * 64))
    for r0 in range(0, n, rows):
        yield r0, min(n, r0 + rows)


# ----------------------------------------------------------------------------------------------- kernel wrappers
def pairwise_distance(x: torch.Tensor, y: torch.Tensor, metric: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gg_pairwise_distance: [n, d] x [m, d] -> [n, m] fp32 on the device."""
    n, d = x.shape
    m, d2 = y.shape
    if d != d2:
        raise ValueError(f"feature dimensions differ: {d} vs {d2}")
    if out is None:
        out = torch.empty(n, m, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().gg_pairwise_distance(_ptr(x), x.stride(0), _ptr(y), y.stride(0), n, m, d, metric, _ptr(out),
                                               out.stride(0), _stream()))
    return out


def row_kth_smallest(dist: torch.Tensor, k: int, want_argmin: bool = False):
    """gg_row_kth_smallest: sorted(dist[i])[k] per row (0-based rank)."""
    n, m = dist.shape
    kth = torch.empty(n, device=dist.device, dtype=torch.float32)
    arg = torch.empty(n, device=dist.device, dtype=torch.int32) if want_argmin else None
    _lib.check(_lib.lib().gg_row_kth_smallest(_ptr(dist), dist.stride(0), n, m, int(k), _ptr(kth), _ptr(arg), _stream()))
    return (kth, arg) if want_argmin else kth


def row_membership(dist: torch.Tensor, col_radius: Optional[torch.Tensor], inclusive: bool, eps: float = 0.0,
                   want=("any", "min", "argmin", "ratio")):
    """gg_row_membership: per-row any / min / argmin / max-ratio against per-column radii."""
    n, m = dist.shape
    dev = dist.device
    out = {
        "any": torch.empty(n, device=dev, dtype=torch.uint8) if "any" in want and col_radius is not None else None,
        "min": torch.empty(n, device=dev, dtype=torch.float32) if "min" in want else None,
        "argmin": torch.empty(n, device=dev, dtype=torch.int32) if "argmin" in want else None,
        "ratio": torch.empty(n, device=dev, dtype=torch.float32) if "ratio" in want and col_radius is not None else None,
    }
    _lib.check(_lib.lib().gg_row_membership(_ptr(dist), dist.stride(0), n, m, _ptr(col_radius), int(inclusive),
                                            float(eps), _ptr(out["any"]), _ptr(out["min"]), _ptr(out["argmin"]),
                                            _ptr(out["ratio"]), _stream()))
    return out


def col_hits(dist: torch.Tensor, row_radius: torch.Tensor, inclusive: bool, hits: torch.Tensor) -> None:
    """gg_col_hits: hits[j] += #{i: dist[i, j] < row_radius[i]} (int32, accumulating)."""
    n, m = dist.shape
    _lib.check(_lib.lib().gg_col_hits(_ptr(dist), dist.stride(0), n, m, _ptr(row_radius), int(inclusive), _ptr(hits),
                                      _stream()))


def standardize_columns(x: torch.Tensor) -> torch.Tensor:
    """gg_standardize_columns: (x - mean) / std per column, constant columns -> 0."""
    n, g = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().gg_standardize_columns(_ptr(x), x.stride(0), n, g, _ptr(out), out.stride(0), _stream()))
    return out


def _kth_neighbour_distance(feats: torch.Tensor, metric: int, rank: int) -> torch.Tensor:
    """Distance from every row to its rank-th neighbour within `feats` (rank 0 = itself), chunked over rows."""
    n = feats.shape[0]
    radii = torch.empty(n, device=feats.device, dtype=torch.float32)
    for r0, r1 in _row_chunks(n, n):
        d = pairwise_distance(feats[r0:r1], feats, metric)
        radii[r0:r1] = row_kth_smallest(d, rank)
    return radii


# ------------------------------------------------------------------------- src/distribution_distances.py:51-142
def compute_pairwise_distance(data_x, data_y=None) -> np.ndarray:
    """L1 distances [N, M] (reference: sklearn pairwise_distances(metric='l1'), :51-66)."""
    dev = _device()
    x = _f32(data_x, dev)
    y = x if data_y is None else _f32(data_y, dev)
    return pairwise_distance(x, y, DIST_L1).cpu().numpy()


def get_kth_value(unsorted, k, axis=-1) -> np.ndarray:
    """k-th smallest value of every row (1-based, :69-83)."""
    a = np.asarray(unsorted) if not isinstance(unsorted, torch.Tensor) else unsorted
    if a.ndim != 2 or axis not in (-1, 1):
        raise ValueError("get_kth_value: 2-D input, last axis only")
    return row_kth_smallest(_f32(a, _device()), int(k) - 1).cpu().numpy()


def compute_nearest_neighbour_distances(input_features, nearest_k) -> np.ndarray:
    """Distance to the nearest_k-th neighbour (:86-99)."""
    return _kth_neighbour_distance(_f32(input_features, _device()), DIST_L1, int(nearest_k)).cpu().numpy()


def compute_prdc(real_features, fake_features, nearest_k) -> dict:
    """Precision, recall, density and coverage of the fake manifold against the real one (:102-142)."""
    dev = _device()
    real, fake = _f32(real_features, dev), _f32(fake_features, dev)
    k = int(nearest_k)
    r_real = _kth_neighbour_distance(real, DIST_L1, k)
    r_fake = _kth_neighbour_distance(fake, DIST_L1, k)
    n, m = real.shape[0], fake.shape[0]
    hits = torch.zeros(m, device=dev, dtype=torch.int32)
    row_any = torch.empty(n, device=dev, dtype=torch.uint8)
    row_min = torch.empty(n, device=dev, dtype=torch.float32)
    for r0, r1 in _row_chunks(n, m):
        d = pairwise_distance(real[r0:r1], fake, DIST_L1)
        res = row_membership(d, r_fake, inclusive=False, want=("any", "min"))
        row_any[r0:r1], row_min[r0:r1] = res["any"], res["min"]
        col_hits(d, r_real[r0:r1], False, hits)
    hits_h = hits.cpu().numpy()
    covered = row_min.cpu().numpy() < r_real.cpu().numpy()
    return dict(precision=(hits_h > 0).mean(), recall=row_any.cpu().numpy().astype(bool).mean(),
                density=(1.0 / float(k)) * hits_h.mean(), coverage=covered.mean())


# --------------------------------------------------------------------------- src/unsupervised_metrics.py:114-345
def batch_pairwise_distances(U, V) -> torch.Tensor:
    """Squared Euclidean distances [len(U), len(V)] (:114-138); returned on the device."""
    dev = _device()
    return pairwise_distance(_f32(U, dev), _f32(V, dev), DIST_SQL2)


class ManifoldEstimator:
    """k-NN hypersphere manifold of `features` (:141-245). `D[:, q]` = squared distance to the nhood_sizes[q]-th
    neighbour. row_batch_size / col_batch_size are accepted for signature compatibility; chunking follows CHUNK_BYTES."""

    def __init__(self, features, row_batch_size=25000, col_batch_size=50000, nhood_sizes: Sequence[int] = (3,),
                 clamp_to_percentile=None, eps=1e-5):
        dev = _device()
        self.nhood_sizes = list(nhood_sizes)
        self.num_nhoods = len(self.nhood_sizes)
        self.eps = eps
        self.row_batch_size, self.col_batch_size = row_batch_size, col_batch_size
        self._ref_features = _f32(features, dev)
        n = self._ref_features.shape[0]
        radii = torch.empty(self.num_nhoods, n, device=dev, dtype=torch.float32)
        for r0, r1 in _row_chunks(n, n):
            d = pairwise_distance(self._ref_features[r0:r1], self._ref_features, DIST_SQL2)
            for q, k in enumerate(self.nhood_sizes):
                radii[q, r0:r1] = row_kth_smallest(d, int(k))
        self.D = np.ascontiguousarray(radii.cpu().numpy().T)
        if clamp_to_percentile is not None:
            max_distances = np.percentile(self.D, clamp_to_percentile, axis=0)
            self.D[self.D > max_distances] = 0
        self._radii_dev = torch.from_numpy(np.ascontiguousarray(self.D.T)).to(dev)

    def evaluate(self, eval_features, return_realism=False, return_neighbors=False):
        dev = self._ref_features.device
        ev = _f32(eval_features, dev)
        num_eval, num_ref = ev.shape[0], self._ref_features.shape[0]
        pred = torch.empty(self.num_nhoods, num_eval, device=dev, dtype=torch.uint8)
        realism = torch.empty(num_eval, device=dev, dtype=torch.float32)
        nearest = torch.empty(num_eval, device=dev, dtype=torch.int32)
        for r0, r1 in _row_chunks(num_eval, num_ref):
            d = pairwise_distance(ev[r0:r1], self._ref_features, DIST_SQL2)
            for q in range(self.num_nhoods):
                res = row_membership(d, self._radii_dev[q], inclusive=True, eps=self.eps,
                                     want=("any", "argmin", "ratio") if q == 0 else ("any",))
                pred[q, r0:r1] = res["any"]
                if q == 0:
                    realism[r0:r1], nearest[r0:r1] = res["ratio"], res["argmin"]
        batch_predictions = np.ascontiguousarray(pred.cpu().numpy().T.astype(np.int32))
        if return_realism and return_neighbors:
            return batch_predictions, realism.cpu().numpy(), nearest.cpu().numpy()
        if return_realism:
            return batch_predictions, realism.cpu().numpy()
        if return_neighbors:
            return batch_predictions, nearest.cpu().numpy()
        return batch_predictions


def knn_precision_recall_features(ref_features, eval_features, nhood_sizes=(3,), row_batch_size=10000,
                                  col_batch_size=50000, num_gpus=1) -> dict:
    """:247-299."""
    ref_manifold = ManifoldEstimator(ref_features, row_batch_size, col_batch_size, nhood_sizes)
    eval_manifold = ManifoldEstimator(eval_features, row_batch_size, col_batch_size, nhood_sizes)
    return dict(precision=ref_manifold.evaluate(eval_features).mean(axis=0),
                recall=eval_manifold.evaluate(ref_features).mean(axis=0))


def get_precision_recall(real_data, fake_data, nb_nn=(10,)):
    """:302-324."""
    state = knn_precision_recall_features(real_data, fake_data, nhood_sizes=list(nb_nn))
    return state["precision"][0], state["recall"][0]


def get_realism_score(real_data, fake_data):
    """:327-345."""
    real_manifold = ManifoldEstimator(real_data, clamp_to_percentile=50)
    _, realism_scores = real_manifold.evaluate(fake_data, return_realism=True)
    return realism_scores


# ------------------------------------------------------------------------------ src/privacy_evaluator.py:9-66
def _neighbour_distances(gen: torch.Tensor, other: torch.Tensor, ranks: Sequence[int]):
    n, m = gen.shape[0], other.shape[0]
    outs = [torch.empty(n, device=gen.device, dtype=torch.float32) for _ in ranks]
    for r0, r1 in _row_chunks(n, m):
        d = pairwise_distance(gen[r0:r1], other, DIST_L2)
        for o, k in zip(outs, ranks):
            o[r0:r1] = row_kth_smallest(d, k)
    return [o.cpu().numpy() for o in outs]


def dcr(real_data, gen_data, test_data, batch_size=128) -> float:
    """Distance to closest record: share of generated rows nearer to a training row than to any test row (:9-32).
    `batch_size` is accepted for compatibility (the reference needs it to bound its [b, N, G] broadcast)."""
    dev = _device()
    real, gen, test = _f32(real_data, dev), _f32(gen_data, dev), _f32(test_data, dev)
    (d_real,) = _neighbour_distances(gen, real, (0,))
    (d_test,) = _neighbour_distances(gen, test, (0,))
    return int((d_real < d_test).sum()) / d_real.shape[0]


def nndr(real_data, gen_data, test_data, batch_size=128) -> float:
    """Nearest-neighbour distance ratio, first / second neighbour, train vs test (:34-66)."""
    dev = _device()
    real, gen, test = _f32(real_data, dev), _f32(gen_data, dev), _f32(test_data, dev)
    r1, r2 = _neighbour_distances(gen, real, (0, 1))
    t1, t2 = _neighbour_distances(gen, test, (0, 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        a, b = r1 / r2, t1 / t2
    return int((a < b).sum()) / a.shape[0]


# ------------------------------------------------------------------------------------ src/corr_score.py:20-120
def upper_diag_list(m_) -> np.ndarray:
    """Strict upper triangle of a square matrix, row by row (:20-40). Indexing only."""
    m = np.asarray(m_)
    return m[np.triu_indices(m.shape[0], k=1)]


def pearson_correlation(x, y) -> np.ndarray:
    """Gene-gene correlation [genes_x, genes_y] of two [samples, genes] matrices (:43-68); 1-D inputs give the scalar
    correlation of the two lists, as np.dot does in the reference."""
    dev = _device()
    one_d = np.ndim(x) == 1
    xa = _f32(np.asarray(x)[:, None] if one_d else x, dev)
    ya = _f32(np.asarray(y)[:, None] if np.ndim(y) == 1 else y, dev)
    if xa.shape[0] != ya.shape[0]:
        raise AssertionError("pearson_correlation: sample counts differ")
    xs, ys = standardize_columns(xa), standardize_columns(ya)
    n, ga = xs.shape
    gb = ys.shape[1]
    out = torch.empty(ga, gb, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().gg_gene_correlation(_ptr(xs), xs.stride(0), _ptr(ys), ys.stride(0), n, ga, gb, _ptr(out),
                                              out.stride(0), _stream()))
    res = out.cpu().numpy()
    return res[0, 0] if one_d else res


def correlations_list(x, y) -> np.ndarray:
    """:91-104."""
    return upper_diag_list(pearson_correlation(x, y))


def gamma_moments(x, y) -> np.ndarray:
    """Six fp64 sums over the gene pairs i < j (count, sum cx, sum cy, sum cx^2, sum cy^2, sum cx*cy) from the fused
    gg_gamma_moments kernel; neither [G, G] correlation matrix is formed."""
    dev = _device()
    xs, ys = standardize_columns(_f32(x, dev)), standardize_columns(_f32(y, dev))
    g = xs.shape[1]
    if ys.shape[1] != g:
        raise ValueError("gamma_coef: the two matrices must have the same genes")
    L = _lib.lib()
    ws = torch.empty(int(L.gg_gamma_moments_workspace_bytes(g)), device=dev, dtype=torch.uint8)
    sums = torch.empty(6, device=dev, dtype=torch.float64)
    _lib.check(L.gg_gamma_moments(_ptr(xs), xs.stride(0), xs.shape[0], _ptr(ys), ys.stride(0), ys.shape[0], g, _ptr(ws),
                                  ws.numel(), _ptr(sums), _stream()))
    return sums.cpu().numpy()


def gamma_coef(x, y) -> float:
    """Gamma(D^X, D^Z): Pearson correlation between the gene-gene distance lists 1 - corr of the two expression
    matrices (:106-120). corr(1 - a, 1 - b) = corr(a, b), evaluated from the six moments."""
    n, sa, sb, saa, sbb, sab = (float(v) for v in gamma_moments(x, y))
    ma, mb = sa / n, sb / n
    va, vb = saa / n - ma * ma, sbb / n - mb * mb
    return (sab / n - ma * mb) / float(np.sqrt(va * vb))


def gamma_coeff_score(x_test, x_gen) -> float:
    """:71-88 (same computation as gamma_coef)."""
    return gamma_coef(x_test, x_gen)
