"""The GEMM epilogue of the hot path (gemmgan_b200/csrc/epilogue.cuh — bias, pre-activation add, LeakyReLU / FiLM
activation, Philox dropout, mask scaling, residual, bf16 / fp32 outputs, accumulation, output-row remap — shared by the
tcgen05 kernel and the CUDA-core check kernel of gemm.cu) checked WITHOUT a GPU through gg_gemm_bf16 with
impl = GG_IMPL_SIMT_F32 on the host emulation (tests/cuda_emu/emu_gemm.cpp), against torch fp32 on the same bf16
operands: what nn.Linear + activation + the torch.cat((x, c), 1) second K segment of the reference compute
(src/conditional_gan_cross_attention_with_film.py:56-72, :129-162, :226). On the B200, tests/test_gpu_gemm.py checks the
tensor-core kernel against the same references AND bit for bit against this check kernel's dropout masks."""
import pytest
import torch

import emu_build
from gemmgan_b200 import _lib, ops


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    L = emu_build.build("gemm", tmp_path_factory.mktemp("cuda_emu"), cudart=True)
    import ctypes as C
    L.gg_gemm_bf16.argtypes = [C.POINTER(_lib.GemmDesc), C.c_void_p]
    return L


@pytest.fixture()
def gemm(emu, monkeypatch):
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setattr(ops, "_stream", lambda: None)

    def run(*a, **kw):
        return ops.gemm(*a, impl=_lib.IMPL_SIMT_F32, **kw)
    return run


def mk(rows, cols, g, ld=None, scale=1.0):
    ld = ld or (cols + 7) // 8 * 8
    buf = torch.zeros(rows, ld, dtype=torch.bfloat16)
    buf[:, :cols] = (torch.randn(rows, cols, generator=g) * scale).bfloat16()
    return buf[:, :cols]


def ref(a, b, a_mn, b_mn):
    af = a.float().t() if a_mn else a.float()
    bf = b.float().t() if b_mn else b.float()
    return af @ bf.t()


def leaky(x, s):
    return torch.where(x > 0, x, s * x)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(100, 40, 64, 0, 0), (33, 70, 17, 0, 1), (26, 100, 20, 1, 1), (52, 31, 100, 1, 0),
                                             (1, 1, 1, 0, 0)])
def test_every_epilogue_feature(gemm, M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    a = mk(K, M, g) if a_mn else mk(M, K, g)
    b = mk(K, N, g) if b_mn else mk(N, K, g)
    bias, pre, mask, res = torch.randn(N, generator=g), mk(M, N, g), mk(M, N, g), torch.randn(M, N, generator=g)
    ob = torch.full((M, N + 10), 7.0, dtype=torch.bfloat16)
    of = torch.full((M, N + 5), 7.0)
    gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, pre=pre, act=_lib.ACT_LEAKY, slope=0.2, mask=mask,
         mask_pos=1.5, mask_neg=-0.5, res=res, alpha=0.5, out_bf16=ob[:, :N], out_f32=of[:, :N])
    want = leaky(0.5 * ref(a, b, a_mn, b_mn) + bias + pre.float(), 0.2)
    want = want * torch.where(mask.float() > 0, 1.5, -0.5) + res
    scale = want.abs().max().item()
    assert (of[:, :N] - want).abs().max().item() <= 1e-4 * scale
    assert (ob[:, :N].float() - want).abs().max().item() <= 5e-3 * scale
    assert (ob[:, N:] == 7.0).all() and (of[:, N:] == 7.0).all()       # nothing outside the [M, N] window
    acc = torch.ones(M, N)
    gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=acc, accum=True)
    assert (acc - (1.0 + ref(a, b, a_mn, b_mn))).abs().max().item() <= 1e-4 * scale


def test_two_segments_row_map_and_film(gemm):
    g = torch.Generator().manual_seed(5)
    Bn, P, S, E, K, K2 = 7, 5, 6, 64, 96, 40
    M = Bn * P
    a, b, a2, b2 = mk(M, K, g), mk(E, K, g), mk(M, K2, g), mk(E, K2, g)
    out = torch.zeros(Bn * S, E, dtype=torch.bfloat16)
    gemm(a, b, a2=a2, b2=b2, out_bf16=out, row_map=(P, S, 1))      # torch.cat((x, c), 1) as a second K segment (:157);
    want = (ref(a, b, 0, 0) + ref(a2, b2, 0, 0)).view(Bn, P, E)   # patch rows land behind each sample's CLS row (:142)
    got = out.view(Bn, S, E)
    assert (got[:, 0] == 0).all()
    assert (got[:, 1:].float() - want).abs().max().item() <= 5e-3 * want.abs().max().item()
    w = mk(128, K, g, scale=2.0)
    of = torch.empty(M, 128)
    gemm(a, w, act=_lib.ACT_FILM, out_f32=of)                      # gamma = tanh(.), beta = clamp(., -5, 5) (:129-134)
    r = ref(a, w, 0, 0)
    want = torch.cat([torch.tanh(r[:, :64]), r[:, 64:].clamp(-5, 5)], 1)
    assert (of - want).abs().max().item() <= 1e-4


def test_dropout_epilogue_statistics_and_determinism(gemm):
    g = torch.Generator().manual_seed(8)
    M, N, K, p = 64, 96, 32, 0.25
    a, b = mk(M, K, g), mk(N, K, g)
    rng = torch.tensor([99, 3], dtype=torch.int64)
    o1, o2, o0 = torch.empty(M, N), torch.empty(M, N), torch.empty(M, N)
    gemm(a, b, drop_p=p, rng=rng, site=4, out_f32=o1)
    gemm(a, b, drop_p=p, rng=rng, site=4, out_f32=o2)
    gemm(a, b, out_f32=o0)
    assert torch.equal(o1, o2)                                      # same (seed, step, site) -> same mask
    kept = o1 != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 0.03
    assert (o1[kept] - o0[kept] / (1 - p)).abs().max().item() <= 1e-5 * o0.abs().max().item()   # inverted dropout
    o3 = torch.empty(M, N)
    gemm(a, b, drop_p=p, rng=rng, site=5, out_f32=o3)
    assert not torch.equal(o1 != 0, o3 != 0)                        # another site draws another mask


def test_argument_errors(emu, gemm):
    g = torch.Generator().manual_seed(1)
    a, b = mk(8, 16, g), mk(8, 16, g)
    with pytest.raises(_lib.GGError):
        gemm(a, b)                                                  # no output
    with pytest.raises(_lib.GGError):
        gemm(a, b, drop_p=0.5, out_f32=torch.empty(8, 8))           # dropout without an rng state
