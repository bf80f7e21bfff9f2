"""One GEMM shape, a few launches (for `ncu --set full`; not a pytest file).

    python tests/gpu_gemm_one.py M N K [bn] [a_mn] [b_mn] [out] [splits]
"""
import sys

import torch

sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops  # noqa: E402

_lib.require_device(0)
M, N, K = (int(x) for x in sys.argv[1:4])
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a_mn = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
b_mn = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
out = sys.argv[7] if len(sys.argv) > 7 else "bf16"
splits = int(sys.argv[8]) if len(sys.argv) > 8 else 0
a = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
b = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if out == "bf16" else None
of = torch.empty(M, N, device="cuda", dtype=torch.float32) if out == "f32" else None
ws = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
bias = torch.randn(N, device="cuda")
for _ in range(3):
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, out_bf16=ob, out_f32=of, workspace=ws, splits=splits, block_n=bn)
torch.cuda.synchronize()
print("ok")
