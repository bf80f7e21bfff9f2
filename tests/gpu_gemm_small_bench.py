"""Tile-width sweep for the one-tile-per-CTA GEMMs of the step (diagnostics; not a pytest file)."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from gemmgan_b200 import _lib  # noqa: E402
from gpu_gemm_bench import bench  # noqa: E402

_lib.require_device(0)
ws = torch.empty(128 << 20, device="cuda", dtype=torch.uint8)
for name, M, N, K, b_mn in [("trunk small", 1024, 256, 256, 0), ("small 3B", 3072, 256, 256, 0), ("dgrad small", 2048, 256, 256, 1),
                            ("text enc", 1024, 256, 768, 0), ("1024x512", 1024, 512, 256, 0), ("film", 1024, 2048, 768, 0)]:
    for bn in (64, 128):
        us, tf, gbs = bench(M, N, K, False, bool(b_mn), "bf16", bn=bn, ws=None, pair=-1)
        print(f"{name:14s} M={M:5d} N={N:5d} K={K:5d} bn={bn:3d} {us:7.2f} us", flush=True)
