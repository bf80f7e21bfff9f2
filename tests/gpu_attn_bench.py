"""Micro-benchmark of the self-attention kernels at BASELINE config 2 / 4 shapes (not a test): forward and
forward+backward of gg_attention_fwd / gg_attention_bwd with and without dropout, CUDA events, L2 flushed between runs.
    python tests/gpu_attn_bench.py"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemmgan_b200 import ops  # noqa: E402


def run(nb, S, p, iters=5, bits=False):
    H, E = 4, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (torch.randn(nb * S, 3 * E, device="cuda", generator=g) * 0.5).bfloat16()
    dout = (torch.randn(nb * S, E, device="cuda", generator=g) * 0.1).bfloat16()
    mask = torch.zeros(nb, S, dtype=torch.uint8, device="cuda")
    rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for name, d in (("fwd", None), ("fwd+bwd", dout)):
        ts = []
        for _ in range(iters + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.attention(qkv, nb, H, S, mask=mask, drop_p=p, rng=rng, site=3, dout=d, precomputed_bits=bits)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name] = sorted(ts[2:])[len(ts[2:]) // 2]
    fl = 4.0 * S * S * E * nb
    print(f"nb={nb} S={S} p={p} bits={int(bits)}: fwd {res['fwd']*1e3:.0f} us ({fl/res['fwd']/1e9:.0f} TFLOP/s), "
          f"bwd {(res['fwd+bwd']-res['fwd'])*1e3:.0f} us ({2.5*fl/(res['fwd+bwd']-res['fwd'])/1e9:.0f} TFLOP/s)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":      # long-kernel sizes, dropout on with precomputed bits
        for S in (129, 145, 161, 193, 225, 257, 289, 320):
            run(768, S, 0.1, bits=True)
        sys.exit(0)
    for nb, S in ((768, 257), (12288, 65), (768, 129)):
        for p, bits in ((0.0, False), (0.1, False), (0.1, True)):
            run(nb, S, p, bits=bits)
