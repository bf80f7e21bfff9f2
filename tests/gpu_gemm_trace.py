"""Pipeline timeline of one CTA of one GEMM launch (diagnostics; not a pytest file).

    python tests/gpu_gemm_trace.py M N K [bn] [pair] [cta]

Prints, relative to the CTA's start (SM clock cycles): when the producer issued each k-block's TMA loads, when
the MMA thread saw each stage full, when it committed each tile, when the first epilogue warp saw each
accumulator full and when it had stored the tile.
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops  # noqa: E402

_lib.require_device(0)
L = _lib.lib()
M, N, K = (int(x) for x in sys.argv[1:4])
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 128
pair = int(sys.argv[5]) if len(sys.argv) > 5 else -1
cta = int(sys.argv[6]) if len(sys.argv) > 6 else 0
do_flush = int(sys.argv[7]) if len(sys.argv) > 7 else 1
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda")
flush = torch.empty(512 << 20, device="cuda", dtype=torch.uint8)
for _ in range(3):
    ops.gemm(a, b, bias=bias, out_bf16=ob, block_n=bn, pair=pair)
if do_flush:
    flush.fill_(1)
torch.cuda.synchronize()
tr = torch.zeros(6 * 512, device="cuda", dtype=torch.int64)
L.gg_gemm_set_trace(C.c_void_p(tr.data_ptr()), cta)
ops.gemm(a, b, bias=bias, out_bf16=ob, block_n=bn, pair=pair)
torch.cuda.synchronize()
L.gg_gemm_set_trace(None, 0)
t = tr.cpu().view(6, 512)
t0 = int(t[5, 0])
names = ["tma_issue(kb)", "stage_full(kb)", "tile_commit", "acc_full", "tile_stored"]
for r, nm in enumerate(names):
    vals = [int(v) - t0 for v in t[r] if int(v) != 0]
    print(f"{nm:16s} n={len(vals):3d}:", " ".join(str(v) for v in vals[:64]))
fine = [int(v) - t0 for v in t[5, 1:97] if int(v) != 0]
print("fine (per chunk of epilogue warp 0, tiles 0-3: start, acc in regs, math done, slot free, staged, store issued):")
for i in range(0, len(fine), 6):
    print("   ", fine[i:i + 6], "deltas", [fine[i + j + 1] - fine[i + j] for j in range(min(5, len(fine) - i - 1))])
