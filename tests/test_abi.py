"""CPU-side checks of the C-ABI library: it builds/loads and exports every symbol include/gemmgan.h declares
(no compute calls — there is no GPU in the build container)."""
import ctypes as C
import os
import re

from gemmgan_b200 import _abi_decl, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "gemmgan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_functions()
    assert "gg_gemm_bf16" in names and "gg_engine_disc_grads" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in gemmgan.h but not exported"
    assert set(_abi_decl.EXPORTS) <= set(names)
    assert L.gg_abi_version() == _lib.ABI_VERSION == 4


def test_struct_layouts_match_header_sizes(tmp_path):
    """ctypes mirrors vs. the real C layout: compile a probe against include/gemmgan.h with gcc."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("gcc not available")
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "gemmgan.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %d %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gg_epilogue), sizeof(gg_gemm_seg),'
        'sizeof(gg_gemm_desc), sizeof(gg_net_buffers), sizeof(gg_model_cfg), offsetof(gg_gemm_desc, epi),'
        'offsetof(gg_net_buffers, off), (int)GG_NSLOTS, sizeof(gg_wgrad_item), offsetof(gg_wgrad_item, bias),'
        'sizeof(gg_colsum_item), sizeof(gg_attn_args), offsetof(gg_gemm_desc, pair), sizeof(gg_enc_layer_params),'
        'offsetof(gg_enc_layer_params, mask_mod), offsetof(gg_enc_layer_params, mean1), sizeof(gg_enc_ffn_bwd_params),'
        'offsetof(gg_enc_ffn_bwd_params, site));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals == [C.sizeof(_lib.Epilogue), C.sizeof(_lib.GemmSeg), C.sizeof(_lib.GemmDesc),
                    C.sizeof(_abi_decl.NetBuffers), C.sizeof(_abi_decl.ModelCfg), _lib.GemmDesc.epi.offset,
                    _abi_decl.NetBuffers.off.offset, _abi_decl.NSLOTS, C.sizeof(_abi_decl.WgradItem),
                    _abi_decl.WgradItem.bias.offset, C.sizeof(_abi_decl.ColsumItem), C.sizeof(_abi_decl.AttnArgs),
                    _lib.GemmDesc.pair.offset, C.sizeof(_abi_decl.EncLayerParams), _abi_decl.EncLayerParams.mask_mod.offset,
                    _abi_decl.EncLayerParams.mean1.offset, C.sizeof(_abi_decl.EncFfnBwdParams),
                    _abi_decl.EncFfnBwdParams.site.offset]


def test_no_cpu_fallback_in_product_path():
    """The product package must not import the oracle or run torch math as a fallback."""
    pkg = os.path.join(ROOT, "gemmgan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn
            # the host emulation of the kernels (tests/cuda_emu) is test infrastructure: the product binds exactly one
            # library, libgemmgan_sm100a.so, and never an emulated build
            assert "cuda_emu" not in text and "_emu.so" not in text and "emu_build" not in text, fn
    for fn in ("vanilla_gan_unconditional.py", "conditional_gan_film.py", "conditional_gan_cross_attention_with_film.py",
               "conditional_gan_cross_attention.py", "conditional_gan_img_transformer.py", "conditional_gan_concat.py",
               "conditional_gan_attention.py",
               "benchmark_generative_model.py", "multi_patch_gan_dataloader.py",
               "multi_patch_multi_token_gan_dataloader.py", "data_loader.py", "benchmark_gan_dataloader.py"):
        text = open(os.path.join(ROOT, fn)).read()
        assert "oracle" not in text.replace("oracle/", ""), fn


def test_label_conditioned_nets_mirror_the_oracle_layout():
    """benchmark_generative_model drop-in (CPU part): same state_dict keys / shapes / initial values as the oracle
    restatement (itself pinned to the reference), and the C-ABI slot table covers every parameter."""
    import torch

    import benchmark_generative_model as bm
    from oracle import restated

    torch.manual_seed(4)
    gen, disc = bm.WGAN_GP_model_benchmark(16, 203, [], [10, 7], [32, 32, 203], [32, 32, 1])
    torch.manual_seed(4)
    o_gen = restated.Net("gen", "label", 203, 16, 0, [32, 32, 203], vocab_sizes=(10, 7))
    o_disc = restated.Net("disc", "label", 203, 16, 0, [32, 32, 1], vocab_sizes=(10, 7))
    for mine, ref in ((gen, o_gen), (disc, o_disc)):
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
        slots = mine.slot_table()
        assert {id(p) for p in slots.values()} == {id(p) for p in mine.parameters()}
        assert slots[_abi_decl.P_EMB0].shape == (10, 128) and slots[_abi_decl.P_EMB1].shape == (7, 128)
    assert gen.input_dims == 16 + 256 and disc.input_dims == 203 + 256
    import pytest
    with pytest.raises(NotImplementedError):
        bm.generator(16, [], [10], [32, 32, 203])   # one variable: 128 != the hard-coded 256 (reference shape error)


def _cfg(variant, B=1024, G=18868, P=8, T=1):
    c = _abi_decl.ModelCfg()
    c.variant = variant
    c.B, c.G, c.L, c.E, c.H = B, G, 256, 256, 256
    c.Dt, c.Dp, c.P, c.T = 768, 1024, P, T
    c.n_layers, c.n_heads, c.ffn = 2, 4, 512
    c.tower_bias = 1
    c.slope, c.dropout_p, c.gp_weight = 0.0, 0.1, 10.0
    c.clip_d, c.clip_g, c.ln_eps = 10.0, 2.0, 1e-5
    c.optimizer = _abi_decl.OPT_RMSPROP
    c.gemm_impl = _lib.IMPL_TCGEN05
    c.seed = 0
    return c


def test_host_only_entry_points_validate_and_report_errors():
    """Error behaviour of the REAL library on a machine without a GPU (host-only entry points; no compute): a negative
    code and a message behind gg_last_error(), never a crash or a silent default; workspace sizes of the BASELINE
    configurations fit one B200 (180 GB) with room for the batch tensors."""
    from gemmgan_b200.models import VARIANT_IDS

    L = _lib.lib()
    _abi_decl.declare(L)
    L.gg_last_error.restype = C.c_char_p
    n = C.c_int64(0)
    sizes = {}
    for name, kw in (("cfg3", dict(B=1024, G=18868, P=8, T=1)), ("cfg2", dict(B=256, G=18868, P=256, T=1)),
                     ("cfg4", dict(B=4096, G=20000, P=64, T=32))):
        c = _cfg(VARIANT_IDS["paper"], **kw)
        assert L.gg_engine_workspace_bytes(C.byref(c), C.byref(n)) == 0, L.gg_last_error()
        sizes[name] = n.value
        assert 0 < n.value < 120 * 2**30 and n.value % 256 == 0, (name, n.value)
    assert sizes["cfg4"] > sizes["cfg3"]
    bad = [(_cfg(99), b"variant"), (_cfg(VARIANT_IDS["paper"], B=0), b"sizes"),
           (_cfg(VARIANT_IDS["paper"], P=400), b"token counts"), (_cfg(VARIANT_IDS["paper"], T=0), b"token counts")]
    c = _cfg(VARIANT_IDS["paper"])
    c.n_heads = 3
    bad.append((c, b"head"))
    c = _cfg(VARIANT_IDS["paper"])
    c.dropout_p = 1.5
    bad.append((c, b"dropout"))
    for c, word in bad:
        rc = L.gg_engine_workspace_bytes(C.byref(c), C.byref(n))
        assert rc == -1, (word, rc)                                  # GG_ERR_ARG
        assert word in L.gg_last_error(), (word, L.gg_last_error())
    assert L.gg_engine_workspace_bytes(None, C.byref(n)) == -1


def test_product_refuses_to_run_without_a_cuda_device():
    """No CPU fallback: without a GPU the library's device check fails and the drop-in trainers raise on construction."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("this is the no-GPU half; the GPU suite covers the other")
    L = _lib.lib()
    assert L.gg_check_device(0) != 0
    with pytest.raises((_lib.GGError, RuntimeError)):
        _lib.require_device(0)
    import vanilla_gan_unconditional as v
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v.WGAN_GP_nocond(input_dims=200, latent_dims=16, vocab_sizes=[], generator_dims=[32, 32, 200],
                         discriminator_dims=[32, 32, 1])
