"""TEST INFRASTRUCTURE ONLY — generates tests/golden/eval_*.npz by running the UNMODIFIED reference evaluation
functions (/root/reference/src via oracle/ref_shim.py) on CPU. Run in the build container:

    python -m oracle.make_eval_golden

The reference's dcr / nndr move their inputs with `.cuda()` (src/privacy_evaluator.py:10-12); there is no GPU in the
build container, so `torch.Tensor.cuda` is made the identity while they run — the arithmetic is the reference's own.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def samples(seed, n_real, n_fake, n_test, g, shift=0.35, fake_scale=1.2):
    """Clustered fp32 profiles: fake = real distribution, slightly shifted and wider, so no metric saturates."""
    r = np.random.RandomState(seed)
    centres = r.randn(4, g) * 1.5
    def draw(n, scale, off):
        return (centres[r.randint(0, 4, n)] + scale * r.randn(n, g) + off).astype(np.float32)
    return draw(n_real, 1.0, 0.0), draw(n_fake, fake_scale, shift), draw(n_test, 1.0, 0.0)


@contextlib.contextmanager
def cuda_is_identity():
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def main():
    dd = ref_shim.load("distribution_distances")
    um = ref_shim.load("unsupervised_metrics")
    pe = ref_shim.load("privacy_evaluator")
    cs = ref_shim.load("corr_score")
    os.makedirs(OUT, exist_ok=True)

    # PRDC (distribution_distances.py:102-142), two neighbourhood sizes, ragged sizes
    for name, (nr, nf, g, k, seed, fs) in {"eval_prdc_a": (60, 50, 37, 5, 1, 1.2),
                                           "eval_prdc_b": (130, 97, 203, 10, 2, 0.8)}.items():
        real, fake, _ = samples(seed, nr, nf, 4, g, shift=0.2, fake_scale=fs)
        res = dd.compute_prdc(real, fake, k)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), real=real, fake=fake, k=k,
            dist=dd.compute_pairwise_distance(real, fake),
            radii_real=dd.compute_nearest_neighbour_distances(real, k),
            radii_fake=dd.compute_nearest_neighbour_distances(fake, k),
            **{key: np.float64(v) for key, v in res.items()})
        print(name, res)

    # k-NN precision / recall and realism (unsupervised_metrics.py:141-345)
    real, fake, _ = samples(3, 64, 48, 4, 20)
    with contextlib.redirect_stdout(io.StringIO()):
        p, r = um.get_precision_recall(torch.tensor(real), torch.tensor(fake), nb_nn=[3])
        est = um.ManifoldEstimator(torch.tensor(real), nhood_sizes=[3])
        pred, realism, nearest = est.evaluate(torch.tensor(fake), return_realism=True, return_neighbors=True)
        realism_clamped = um.get_realism_score(torch.tensor(real), torch.tensor(fake))
    np.savez_compressed(os.path.join(OUT, "eval_knn_pr.npz"), real=real, fake=fake, k=3, precision=np.float64(p),
                        recall=np.float64(r), radii=est.D, pred=pred, realism=realism, nearest=nearest,
                        realism_clamped=realism_clamped,
                        sqdist=um.batch_pairwise_distances(torch.tensor(fake), torch.tensor(real)).numpy())
    print("eval_knn_pr", p, r)

    # DCR / NNDR (privacy_evaluator.py:9-66); 150 generated rows = one full reference batch of 128 plus a tail
    real, fake, test = samples(4, 70, 150, 40, 33, shift=0.1)
    fake[:20] = real[:20] + 0.05 * np.random.RandomState(5).randn(20, 33).astype(np.float32)  # near copies
    with cuda_is_identity():
        d = pe.dcr(real, fake, test)
        n = pe.nndr(real, fake, test)
    np.savez_compressed(os.path.join(OUT, "eval_privacy.npz"), real=real, fake=fake, test=test, dcr=np.float64(d),
                        nndr=np.float64(n))
    print("eval_privacy", d, n)

    # gene-gene Pearson correlation and the gamma coefficient (corr_score.py:43-120); gene 7 constant in x
    r = np.random.RandomState(6)
    mix = r.randn(12, 50)
    x = (r.randn(40, 12) @ mix + 0.5 * r.randn(40, 50)).astype(np.float32)
    y = (r.randn(30, 12) @ mix + 0.9 * r.randn(30, 50)).astype(np.float32)
    x[:, 7] = 1.25
    y2 = (x + r.randn(40, 50)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        corr = cs.pearson_correlation(x, y2)
        gamma = cs.gamma_coef(x, y)
        gamma_score = cs.gamma_coeff_score(x, y)
        clist = cs.correlations_list(x, x)
    np.savez_compressed(os.path.join(OUT, "eval_gamma.npz"), x=x, y=y, y2=y2, corr=corr, gamma=np.float64(gamma),
                        gamma_score=np.float64(gamma_score), corr_list=clist)
    print("eval_gamma", gamma, gamma_score)


if __name__ == "__main__":
    main()
