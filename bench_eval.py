"""Evaluation-metric kernels timed on one B200 next to the reference's CPU functions.

    python bench_eval.py [--n 2000] [--m 2000] [--genes 18868] [--iters 5] [--cpu-rows 200]

One JSON line per kernel / metric (CUDA events on the launching stream, inputs resident in HBM, warm-up 2):
  pairwise_l1 / pairwise_l2   gg_pairwise_distance on [n, G] x [m, G]; `achieved` = n*m*G element pairs per second
                              expressed as fp32 instruction TFLOP/s (L1: subtract + |.|-accumulate = 2 per pair,
                              L2: subtract + FMA = 3 flops per pair) against the CUDA-core peak
                              148 SMs x 128 lanes x 2 x SM clock (nominal 74.4 TFLOP/s at 1965 MHz);
  row_kth                     gg_row_kth_smallest rank 10 on the [n, m] matrix; bytes = 11 passes x 4*n*m (L2-resident);
  gamma                       gg_gamma_moments on two standardised [n, G] matrices: G^2*(2n)/2 FMAs;
  prdc / dcr                  the whole reference-named call (host buffers in, dict / float out) = the e2e figure;
  cpu                         oracle/evalmetrics_ref.py (the reference's algorithm in numpy) on --cpu-rows rows of the
                              same problem, scaled to the full row count and labelled as such.
Nothing here is a `bench.py` line: the headline metric of BASELINE.json is the training step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timed(fn, iters, torch):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--m", type=int, default=2000)
    ap.add_argument("--genes", type=int, default=18868)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-rows", type=int, default=200)
    args = ap.parse_args()
    import numpy as np
    import torch

    from gemmgan_b200 import _lib
    from gemmgan_b200 import evalmetrics as em

    _lib.require_device(0)
    dev = torch.device("cuda", 0)
    n, m, G = args.n, args.m, args.genes
    gen = torch.Generator(device=dev).manual_seed(42)
    real = torch.randn(n, G, device=dev, generator=gen)
    fake = torch.randn(m, G, device=dev, generator=gen) * 1.1 + 0.05
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    lines = []

    def emit(**kw):
        lines.append(kw)
        print(json.dumps(kw), flush=True)

    out = torch.empty(n, m, device=dev)
    for name, metric, flops in (("pairwise_l1", em.DIST_L1, 2), ("pairwise_l2", em.DIST_L2, 3)):
        t = timed(lambda: em.pairwise_distance(real, fake, metric, out), args.iters, torch)
        tf = flops * n * m * G / t / 1e12
        emit(kernel=name, n=n, m=m, genes=G, ms=t * 1e3, achieved=tf, peak=peak, unit="TFLOP/s (fp32 CUDA cores)",
             frac=tf / peak)
    t = timed(lambda: em.row_kth_smallest(out, 10), args.iters, torch)
    emit(kernel="row_kth", n=n, m=m, rank=10, ms=t * 1e3, achieved=11 * 4 * n * m / t / 1e9, unit="GB/s (L2-resident passes)")
    xs, ys = em.standardize_columns(real), em.standardize_columns(fake[: min(m, n)])
    ws = torch.empty(int(_lib.lib().gg_gamma_moments_workspace_bytes(G)), device=dev, dtype=torch.uint8)
    sums = torch.empty(6, device=dev, dtype=torch.float64)
    import ctypes as C

    def gamma():
        _lib.check(_lib.lib().gg_gamma_moments(C.c_void_p(xs.data_ptr()), xs.stride(0), xs.shape[0],
                                               C.c_void_p(ys.data_ptr()), ys.stride(0), ys.shape[0], G,
                                               C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(sums.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    t = timed(gamma, args.iters, torch)
    tf = 2 * (G * (G + 64) / 2) * (xs.shape[0] + ys.shape[0]) / t / 1e12
    emit(kernel="gamma_moments", genes=G, nx=xs.shape[0], ny=ys.shape[0], ms=t * 1e3, achieved=tf, peak=peak,
         unit="TFLOP/s (fp32 CUDA cores)", frac=tf / peak)

    # whole calls, host buffers in (the e2e figure) vs the reference's algorithm on the host cores
    real_h, fake_h = real.cpu().numpy(), fake.cpu().numpy()
    for name, call in (("compute_prdc", lambda: em.compute_prdc(real_h, fake_h, 10)),
                       ("dcr", lambda: em.dcr(real_h, fake_h, real_h[: max(2, n // 4)])),
                       ("gamma_coef", lambda: em.gamma_coef(real_h, fake_h))):
        call()
        t0 = time.perf_counter()
        call()
        emit(call=name, n=n, m=m, genes=G, e2e_s=time.perf_counter() - t0, h2d_bytes=int(real_h.nbytes + fake_h.nbytes))
    from oracle import evalmetrics_ref as ref  # CPU baseline leg only

    r = min(args.cpu_rows, n)
    t0 = time.perf_counter()
    ref.compute_pairwise_distance(real_h[:r], fake_h)
    dt = time.perf_counter() - t0
    emit(cpu_baseline="compute_pairwise_distance (numpy fp64, 1 thread)", rows=r, s=dt,
         scaled_to_full_rows_s=dt * n / r, kind="port", cores=1, sample=f"{r} of {n} rows x {m} x {G}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
