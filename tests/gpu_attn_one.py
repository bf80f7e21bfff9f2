"""One forward + backward of the long self-attention kernels at BASELINE config 2's shape (768 sequences x 257 tokens, 4 heads
of 64, dropout 0.1 with the precomputed mask bits) for ncu (not a test).   python tests/gpu_attn_one.py [S] [nb]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemmgan_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 257
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 768
H, E = 4, 256
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(nb * S, 3 * E, device="cuda", generator=g) * 0.5).bfloat16()
dout = (torch.randn(nb * S, E, device="cuda", generator=g) * 0.1).bfloat16()
rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.attention(qkv, nb, H, S, drop_p=0.1, rng=rng, site=3, dout=dout, precomputed_bits=True)
torch.cuda.synchronize()
print("done")
