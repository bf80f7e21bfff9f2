"""Diagnostic sweep of gg_gemm_bf16 on a real B200 (not a pytest file; run by hand under gpurun).

Usage: python tests/gpu_gemm_probe.py <a_mn> <b_mn>      # one operand-major combination per process
Each case prints max|err| / max|ref| for the tcgen05 path and for the CUDA-core check path.
"""
import sys

import torch

sys.path.insert(0, ".")
from gemmgan_b200 import _lib, ops  # noqa: E402


def run_case(a_mn, b_mn, M, N, K, bn=0, splits=0, K2=0, ws=None, impl=_lib.IMPL_TCGEN05):
    g = torch.Generator(device="cuda").manual_seed(1234 + M + 3 * N + 7 * K)
    def mk(rows, cols):
        ld = (cols + 7) // 8 * 8
        buf = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
        buf[:, :cols] = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
        return buf[:, :cols]
    a = mk(K, M) if a_mn else mk(M, K)
    b = mk(K, N) if b_mn else mk(N, K)
    af = (a.float().t() if a_mn else a.float())
    bf = (b.float().t() if b_mn else b.float())
    ref = af @ bf.t()
    a2 = b2 = None
    if K2:
        a2 = mk(K2, M) if a_mn else mk(M, K2)
        b2 = mk(K2, N) if b_mn else mk(N, K2)
        ref = ref + (a2.float().t() if a_mn else a2.float()) @ (b2.float().t() if b_mn else b2.float()).t()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, a2=a2, b2=b2, out_f32=out, workspace=ws, impl=impl,
             splits=splits, block_n=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    return err


def main():
    a_mn, b_mn = int(sys.argv[1]), int(sys.argv[2])
    torch.backends.cuda.matmul.allow_tf32 = False
    _lib.require_device(0)
    ws = torch.empty(64 << 20, device="cuda", dtype=torch.uint8)
    cases = [
        # M, N, K, bn, splits, K2
        (128, 128, 64, 128, 0, 0),
        (128, 128, 256, 128, 0, 0),
        (256, 256, 512, 128, 0, 0),
        (128, 64, 128, 64, 0, 0),
        (128, 256, 128, 256, 0, 0),
        (200, 136, 104, 128, 0, 0),     # ragged everything (multiples of 8)
        (1024, 256, 1000, 128, 4, 0),   # split-K
        (256, 512, 320, 128, 0, 192),   # two K segments
        (384, 1000, 256, 256, 0, 0),
        (100, 72, 40, 64, 0, 0),
        (4096, 1024, 256, 128, 0, 0),   # 256 tiles > 148 SMs: persistent loop + TMEM double buffering
        (1024, 18868, 256, 128, 0, 0),  # generator output shape, ragged N
        (2048, 1280, 192, 256, 0, 0),   # BN=256, 80 tiles
        (27648, 512, 256, 128, 0, 0),   # tower FFN shape
        (512, 256, 9000, 128, 9, 0),    # long K split 9 ways
    ]
    ok = True
    for (M, N, K, bn, splits, K2) in cases:
        for impl, name in ((_lib.IMPL_SIMT_F32, "simt"), (_lib.IMPL_TCGEN05, "tcgen05")):
            try:
                err = run_case(a_mn, b_mn, M, N, K, bn, splits if impl == 0 else 0, K2, ws, impl)
                status = "OK " if err < 2e-3 else "BAD"
                ok &= err < 2e-3
            except Exception as ex:  # noqa: BLE001
                err, status = float("nan"), f"EXC {ex}"
                ok = False
            print(f"a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} K2={K2} bn={bn} splits={splits} {name:8s} rel_err={err:.3e} {status}", flush=True)
    print("PROBE", "PASS" if ok else "FAIL", a_mn, b_mn)


if __name__ == "__main__":
    main()
