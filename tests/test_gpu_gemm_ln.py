"""gg_gemm_layernorm (csrc/gemm_ln.cu): z = res + dropout(a w^T + bias), out = LayerNorm(z) in one tcgen05
kernel, against fp32 torch math on the same bf16 operands; the dropout keep mask is regenerated on the host by the numpy
port of csrc/philox.cuh (tests/test_gpu_enc_layer.py::philox_keep), so the dropout-on configuration is compared exactly."""
import ctypes as C

import numpy as np
import pytest
import torch

from gemmgan_b200 import _lib
from test_gpu_enc_layer import philox_keep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,K,p,bias,beta", [(1000, 256, 0.0, True, True), (16640, 512, 0.1, True, True),
                                               (333, 256, 0.1, False, False), (128, 512, 0.0, True, False),
                                               (65 * 300, 256, 0.1, True, True)])
def test_gemm_layernorm_matches_torch(rows, K, p, bias, beta):
    _lib.require_device(0)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(rows + K)
    r = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale
    a, w = r(rows, K).bfloat16(), r(256, K, scale=K ** -0.5).bfloat16()
    res = r(rows, 256).bfloat16()
    bvec = r(256, scale=0.1) if bias else None
    gamma = 1 + 0.1 * r(256)
    bet = 0.1 * r(256) if beta else None
    seed, step, site = 0xABC123, 9, 25
    rng = torch.tensor([seed, step], dtype=torch.int64, device="cuda")
    z = torch.full((rows, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = torch.full_like(z, float("nan"))
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    ptr = lambda t: None if t is None else t.data_ptr()
    _lib.check(L.gg_gemm_layernorm(a.data_ptr(), K, w.data_ptr(), K, K, ptr(bvec), res.data_ptr(), gamma.data_ptr(), ptr(bet),
                                   z.data_ptr(), out.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, 1e-5, p,
                                   rng.data_ptr(), site, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    y = a.float() @ w.float().t() + (0 if bvec is None else bvec)
    if p > 0:
        idx = np.arange(rows * 256, dtype=np.int64).reshape(rows, 256)
        keep = torch.from_numpy(philox_keep(seed, step, site, idx, p)).cuda().float() / (1 - p)
        y = y * keep
    zf = res.float() + y
    mu, var = zf.mean(1), zf.var(1, unbiased=False)
    want = (zf - mu[:, None]) * torch.rsqrt(var + 1e-5)[:, None] * gamma + (0 if bet is None else bet)
    scale = zf.abs().max().item()
    assert (z.float() - zf).abs().max().item() <= 1e-2 * scale
    assert torch.allclose(mean, mu, atol=2e-3 * scale) and torch.allclose(rstd, torch.rsqrt(var + 1e-5), rtol=1e-2)
    assert (out.float() - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    if p > 0:   # exact zeros of the dropped projection survive as z == res (a mask error would show here, not in a tolerance)
        dropped = keep == 0
        assert torch.equal(z[dropped], res[dropped])
