"""Per-role clock64() trace of one tile of the fused encoder-layer kernel (diagnostics; not a pytest file):
    python tests/gpu_enc_layer_trace.py [nb] [S] [drop_p] [save_rows]
Prints the stamps of CTA 0's first tile relative to the kernel start and the kernel's CUDA-event duration."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_enc_layer as T  # noqa: E402
from gemmgan_b200 import _lib  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 9
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
save = int(sys.argv[4]) if len(sys.argv) > 4 else -1
L = _lib.lib()
torch.manual_seed(1)
x = torch.randn(nb * S, T.E, device="cuda").bfloat16()
W = T.make_weights(True)
for _ in range(3):
    T.run_kernel(x, W, None, nb, S, p, 1, 2, 8, save)
buf = torch.zeros(3, 64, dtype=torch.int64, device="cuda")
L.gg_enc_layer_set_trace(C.c_void_p(buf.data_ptr()))
T.run_kernel(x, W, None, nb, S, p, 1, 2, 8, save)
L.gg_enc_layer_set_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = {0: "producer (tile start, then one stamp per weight stage when its slot is free)",
         1: "mma (start, x_full, [head: begin, issued] x4, ao_full, outproj issued, x1_full, f1a issued, f1b issued, "
            "ha_full, f2a issued, hb_full, f2b issued)",
         2: "epilogue warp 0 ([head: wait, acc_full, staged] x4, attention done, wait, acc2_full, [LN: ldtm, sweep1, fence, bar, sweep2, fence], LN1 done, f1a_full, "
            "ha done, f1b_full+f2a_done, hb done, out_full, LN2 done)"}
for r in range(3):
    row = [int(v) - t0 for v in t[r] if v > 0]
    print(names[r])
    print("  ", row)
import os
if os.environ.get("EL_TRACE_ONLY"):
    sys.exit(0)
# timing of the launch alone (events)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    e0.record()
    # run_kernel allocates outputs; time only a bare relaunch through the same params is not exposed, so this
    # includes ~10 small fill kernels; the ncu launch list gives the kernel alone
    T.run_kernel(x, W, None, nb, S, p, 1, 2, 8, save)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("run_kernel (incl. output fills) ms:", sorted(ts))
