"""gg_film_patch_encode (csrc/film_patch.cu): FiLM modulation in the A-operand path of the patch-encoder GEMM, bias,
CLS rows and the replica copies in one kernel, against torch on the same bf16 operands
(src/conditional_gan_cross_attention_with_film.py:129-142), and the engine with / without it."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

from gemmgan_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("B,P,R,Dp,bias,save", [(1024, 8, 3, 1024, True, True), (37, 5, 1, 128, False, True),
                                                (256, 64, 2, 1024, True, False), (3, 1, 3, 64, True, True)])
def test_film_patch_encode_matches_torch(B, P, R, Dp, bias, save):
    _lib.require_device(0)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(B + P)
    r = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale
    patches = r(B * P, Dp).bfloat16()
    gb = torch.cat((torch.tanh(r(B, Dp)), r(B, Dp).clamp(-5, 5)), dim=1).contiguous()
    w = r(256, Dp, scale=Dp ** -0.5).bfloat16()
    bvec = r(256) if bias else None
    cls = r(256, scale=0.02)
    S = P + 1
    x0 = torch.full((R * B * S, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    mod = torch.full((B * P, Dp), float("nan"), device="cuda", dtype=torch.bfloat16) if save else None
    rc = L.gg_film_patch_encode(patches.data_ptr(), gb.data_ptr(), w.data_ptr(), w.stride(0),
                                None if bvec is None else bvec.data_ptr(), cls.data_ptr(), x0.data_ptr(),
                                None if mod is None else mod.data_ptr(), B, P, R, Dp,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    torch.cuda.synchronize()
    gamma, beta = gb[:, :Dp].repeat_interleave(P, 0), gb[:, Dp:].repeat_interleave(P, 0)
    want_mod = torch.addcmul(beta, gamma, patches.float()).bfloat16()          # fmaf then one rounding, as the kernel
    if save:
        assert torch.equal(mod, want_mod)
    pe = want_mod.float() @ w.float().t() + (0 if bvec is None else bvec)
    want = torch.cat((cls.bfloat16().float().expand(B, 1, 256), pe.view(B, P, 256)), dim=1)
    got = x0.float().view(R, B, S, 256)
    for rep in range(R):
        assert torch.equal(got[rep, :, 0], want[:, 0]), "CLS rows"
        err = (got[rep] - want).abs().max().item() / want.abs().max().item()
        assert err < 1e-2, (rep, err)
    assert torch.equal(got[0], got[R - 1])


def test_engine_with_and_without_the_fused_film_prologue():
    """The same training steps through the engine with film_patch.cu and with the three separate launches
    (GEMMGAN_FILM_FUSED=0; the switch is read when an engine is created, hence one process per setting): the dropout
    streams are the same, so the losses agree to the summation order of one GEMM."""
    code = r"""
import sys, json, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import bench
w = dict(bench.WORKLOADS["cfg3"]); w["B"] = 128; w["G"] = 2000
t = bench.build_trainer(w, "adam")
batch = bench.make_batch(w, seed=7, device=torch.device("cuda"))
torch.manual_seed(5)
out = []
for _ in range(3):
    t.train(*batch)
    out.append([float(x) for x in t.d_batch_loss] + [float(x) for x in t.g_batch_loss])
print("LOSSES" + json.dumps(out))
""" % (ROOT, ROOT)
    res = {}
    for flag in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GEMMGAN_FILM_FUSED=flag), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        import json
        res[flag] = json.loads([l for l in r.stdout.splitlines() if l.startswith("LOSSES")][0][6:])
    a, b = torch.tensor(res["1"]), torch.tensor(res["0"])
    assert torch.allclose(a, b, rtol=2e-2, atol=2e-3), (a, b)
