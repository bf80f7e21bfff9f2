"""gg_xw_f32 (csrc/xw_f32.cu): critic layer 1 on the gradient penalty's two fp32 gene matrices read in place — fp32 -> bf16
on chip, 256-row CTAs, the (tile, k-block) sequence cut into equal shares (stream-K) with a deterministic fix-up of the
tiles two or three shares touched — against torch fp32 math on the bf16-rounded operands (the rounding the kernel
applies), with row / K tails, fewer workers than SMs (small workspace) and single-tile problems."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r"""
import sys, torch
sys.path.insert(0, {root!r})
from gemmgan_b200 import ops
torch.manual_seed(0)
for B, K, ws in {cases!r}:
    x0 = torch.randn(B, K, device="cuda")
    x1 = torch.randn(B, K, device="cuda") * 0.5 + 0.1
    w = (torch.randn(256, (K + 7) // 8 * 8, device="cuda") * 0.05).bfloat16()[:, :K]   # pitch: a multiple of 8 (TMA)
    got = ops.xw_f32(x0, x1, w, workspace_mb=ws)
    torch.cuda.synchronize()
    want = torch.cat((x0, x1)).bfloat16().float() @ w.float().t()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5 * max(K, 1000) ** 0.5, (B, K, ws, err)
    print("ok", B, K, ws, err)
"""


def test_xw_f32_matches_torch():
    """In its own process and under a timeout: a protocol bug in a persistent kernel is a hang, not a wrong number.
    Cases: (rows per tensor, K, workspace MB) — 1 MB of workspace = 2 workers, 16 MB = 32 workers."""
    cases = [(128, 64, 96), (512, 256, 96), (300, 1000, 96), (1024, 20000, 96), (4096, 18868, 96), (5000, 2052, 16),
             (777, 4100, 1), (16384, 2000, 96), (256, 128, 1)]
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-c", CASE.format(root=ROOT, cases=cases)], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == len(cases)
