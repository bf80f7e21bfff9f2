"""Micro-benchmarks of the non-GEMM kernels on the cfg3 shapes (not a pytest file; run under gpurun).
Device time per launch: ITERS launches captured in one CUDA graph, replay timed with CUDA events."""
import sys

import torch

sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops  # noqa: E402

ITERS = 10


def timed(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(ITERS):
            fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / ITERS)
    return sorted(ts)[len(ts) // 2]


def main():
    _lib.require_device(0)
    H, hd = 4, 64
    E = H * hd
    rng = torch.tensor([1, 1], device="cuda", dtype=torch.int64)
    for nb, L in [(3072, 9), (2048, 9), (1024, 9), (4096 * 3, 65)]:
        if nb * L * 3 * E * 2 > 8e9:
            continue
        qkv = torch.randn(nb * L, 3 * E, device="cuda").to(torch.bfloat16)
        dout = torch.randn(nb * L, E, device="cuda").to(torch.bfloat16)
        for p in (0.0, 0.1):
            us_f = timed(lambda: ops.attention(qkv, nb, H, L, drop_p=p, rng=rng))
            us_fb = timed(lambda: ops.attention(qkv, nb, H, L, drop_p=p, rng=rng, dout=dout))
            byts_f = qkv.numel() * 2 + dout.numel() * 2
            print(f"attention nb={nb:6d} L={L:3d} p={p}: fwd {us_f:8.1f} us ({byts_f / us_f / 1e3:7.1f} GB/s)   "
                  f"bwd {us_fb - us_f:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
