"""k-block rate of the tcgen05 GEMM for K-major against MN-major operands (not a test): M = 128, N = 128, one CTA,
K = 64 * kb with no K split -> time / kb.   python tests/gpu_kblock_rate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemmgan_b200 import ops  # noqa: E402


def rate(mn, M=128, N=128, kb=2048, tiles=1, block_n=128, l2=False):
    K = 64 * kb
    g = torch.Generator(device="cuda").manual_seed(0)
    if mn:
        a = torch.randn(K, M * tiles, device="cuda", generator=g).bfloat16()
        b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    else:
        a = torch.randn(M * tiles, K, device="cuda", generator=g).bfloat16()
        b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    out = torch.empty(M * tiles, N, device="cuda", dtype=torch.float32)
    stage_kb = (128 * 64 + block_n * 64) * 2 / 1024
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, b, a_mn=mn, b_mn=mn, out_f32=out, splits=1, block_n=block_n, light=-1, pair=-1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"{'MN' if mn else 'K '}-major  tiles={tiles:3d}  N={N} block_n={block_n} kb={kb}: {t * 1e3:8.1f} us  -> "
          f"{t * 1e6 / kb:7.1f} ns per {stage_kb:.0f} KB k-block per CTA ({stage_kb * 1024 / (t * 1e6 / kb):5.1f} GB/s per SM)")


for tiles in (1, 148):
    for mn in (False, True):
        rate(mn, tiles=tiles)
# ring depth: 128-wide tiles run 4 stages of 32 KB, 256-wide tiles 3 stages of 48 KB; operands from L2 (short K, many
# repeats of the same 1 MB) against operands from HBM (above)
rate(False, N=256, block_n=256, tiles=148)
rate(False, kb=64, tiles=1)
rate(False, N=256, block_n=256, kb=64, tiles=1)
