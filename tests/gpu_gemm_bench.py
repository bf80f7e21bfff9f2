"""Micro-benchmark of gg_gemm_bf16 on the shapes the cfg3 training step uses (not a pytest file).

Prints per shape: time (CUDA events, L2 flushed between iterations), TFLOP/s, GB/s of algorithmic traffic.
"""
import sys

import torch

sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops  # noqa: E402


def bench(M, N, K, a_mn=False, b_mn=False, out="bf16", bias=True, bn=0, splits=0, iters=20, act=0, drop=0.0,
          mask=False, res=False, ws=None, rng=None, flush=None, light=0, pair=-1):
    """Device time per launch: `iters` launches over 4 rotating operand sets are captured in one CUDA graph and
    the replay is timed with CUDA events (no host launch overhead in the number; operands are L2-warm at best
    every 4th launch, as inside the training step)."""
    dev = "cuda"
    sets = []
    for _ in range(4):
        a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if out == "bf16" else None
        of = torch.empty(M, N, device=dev, dtype=torch.float32) if out == "f32" else None
        bs = torch.randn(N, device=dev) if bias else None
        mk = torch.randn(M, N, device=dev).to(torch.bfloat16) if mask else None
        rs = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
        sets.append((a, b, dict(a_mn=a_mn, b_mn=b_mn, bias=bs, act=act, drop_p=drop, rng=rng, mask=mk, res=rs,
                                out_bf16=ob, out_f32=of, workspace=ws, splits=splits, block_n=bn, light=light, pair=pair)))
    for a, b, kw in sets:
        ops.gemm(a, b, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            a, b, kw = sets[i % 4]
            ops.gemm(a, b, **kw)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    ts.sort()
    us = ts[len(ts) // 2]
    flops = 2.0 * M * N * K
    byts = 2.0 * (M * K + N * K) + M * N * (2 if out == "bf16" else 4) + (M * N * 2 if mask else 0) + (M * N * 2 if res else 0)
    return us, flops / us / 1e6, byts / us / 1e3


def main():
    _lib.require_device(0)
    ws = torch.empty(128 << 20, device="cuda", dtype=torch.uint8)
    rng = torch.tensor([1, 1], device="cuda", dtype=torch.int64)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    cases = [
        # name, M, N, K, a_mn, b_mn, out, extras
        ("ffn1 fwd 3B*S (drop)", 27648, 512, 256, 0, 0, "bf16", dict(act=1, drop=0.1)),
        ("ffn1 fwd 3B*S", 27648, 512, 256, 0, 0, "bf16", dict(act=1)),
        ("ffn1 fwd nobias", 27648, 512, 256, 0, 0, "bf16", dict(bias=False)),
        ("qkv fwd 3B*S", 27648, 768, 256, 0, 0, "bf16", {}),
        ("out fwd 3B*S", 27648, 256, 256, 0, 0, "bf16", {}),
        ("ffn2 fwd 3B*S", 27648, 256, 512, 0, 0, "bf16", {}),
        ("patch enc", 8192, 256, 1024, 0, 0, "bf16", {}),
        ("dgrad ffn2 (mask)", 18432, 512, 256, 0, 1, "bf16", dict(bias=False, mask=True)),
        ("dgrad ffn1 (res)", 18432, 256, 512, 0, 1, "bf16", dict(bias=False, res=True)),
        ("wgrad ffn1", 512, 256, 18432, 1, 1, "f32", dict(bias=False)),
        ("wgrad qkv", 768, 256, 18432, 1, 1, "f32", dict(bias=False)),
        ("wgrad out", 256, 256, 18432, 1, 1, "f32", dict(bias=False)),
        ("gen final", 1024, 18872, 256, 0, 0, "bf16", {}),
        ("critic L1 [2B,G]", 2048, 256, 18872, 0, 0, "f32", dict(bias=False)),
        ("dW1x", 256, 18872, 2048, 1, 1, "f32", dict(bias=False)),
        ("gram", 256, 256, 18872, 0, 0, "f32", dict(bias=False)),
        ("trunk small", 1024, 256, 256, 0, 0, "bf16", {}),
        ("trunk small 3B", 3072, 256, 256, 0, 0, "bf16", {}),
        ("square 4096", 4096, 4096, 4096, 0, 0, "bf16", dict(bias=False)),
        ("square 8192", 8192, 8192, 8192, 0, 0, "bf16", dict(bias=False)),
    ]
    for name, M, N, K, a_mn, b_mn, out, ex in cases:
        for bn, pair in ((128, -1), (256, -1), (256, 1)):
            if N < 256 and bn == 256:
                continue
            us, tf, gbs = bench(M, N, K, bool(a_mn), bool(b_mn), out, bn=bn, ws=ws, rng=rng, flush=flush, pair=pair, **ex)
            print(f"{name:24s} M={M:6d} N={N:6d} K={K:6d} bn={bn:3d} pair={pair:2d} {us:8.1f} us  {tf:8.1f} TFLOP/s  {gbs:8.1f} GB/s",
                  flush=True)


if __name__ == "__main__":
    main()
