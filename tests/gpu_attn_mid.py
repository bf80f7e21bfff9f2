import sys, torch
sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops
_lib.require_device(0)
H, hd, nb, L = 4, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 65
E = H * hd
qkv = torch.randn(nb * L, 3 * E, device="cuda").to(torch.bfloat16)
dout = torch.randn(nb * L, E, device="cuda").to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, nb, H, L, dout=dout)
torch.cuda.synchronize()
print("ok")
