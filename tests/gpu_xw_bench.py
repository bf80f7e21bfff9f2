"""Timing of gg_xw_f32 alone at BASELINE config 5 (B = 16384 rows per tensor, G = 20000), L2 flushed between runs, and a
torch.profiler kernel table of one gg_engine_gp_step (not a test).   python tests/gpu_xw_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemmgan_b200 import ops  # noqa: E402

B, K = 16384, 20000
x0 = torch.randn(B, K, device="cuda")
x1 = torch.randn(B, K, device="cuda")
w = (torch.randn(256, K, device="cuda") * 0.05).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.xw_f32(x0, x1, w, workspace_mb=96)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
t = sorted(ts[2:])[len(ts[2:]) // 2]
print(f"cluster={os.environ.get('GEMMGAN_XW_CLUSTER', '4')}: xw_f32 {t * 1e3:.0f} us, {2 * B * K * 4 / t / 1e6:.0f} GB/s of fp32 input")
