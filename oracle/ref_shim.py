"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference scripts from /root/reference/src.

Used in the build container to (a) validate oracle/restated.py against the real reference and
(b) generate the golden vectors under tests/golden/ (oracle/make_golden.py). The reference tree
does not exist on the GPU box, so nothing under tests -m gpu, smoke() or bench.py imports this
file at run time there; those use oracle/restated.py and the committed fixtures.

The reference's model scripts star-import evaluation modules that pull third-party packages
missing from this image (matplotlib, lightgbm, catboost, ot, umap, seaborn, torch_geometric and
two repo-external modules). None of them is touched by WGAN_GP.train(); they are replaced by
inert stubs before the import.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import tempfile
from unittest.mock import MagicMock

REF_SRC = os.environ.get("GEMMGAN_REFERENCE_SRC", "/root/reference/src")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "lightgbm", "catboost", "ot", "umap",
    "seaborn", "torch_geometric", "torch_geometric.nn", "rnaseq_contrastive_model", "contrastive_model",
]


def available() -> bool:
    return os.path.isdir(REF_SRC)


_loaded = {}        # reference scripts imported so far, by name
_ref_modules = {}   # every module that came from REF_SRC (kept out of sys.modules between loads)


def _from_reference(mod) -> bool:
    f = getattr(mod, "__file__", None)
    return bool(f) and os.path.abspath(f).startswith(os.path.abspath(REF_SRC) + os.sep)


def load(module_name: str):
    """Imports one reference script (e.g. 'conditional_gan_cross_attention_with_film').

    The drop-in modules at the repo root carry the reference's file names on purpose, and the reference scripts
    import each other by bare name, so both sets cannot sit in sys.modules together: while a reference script is
    being imported the same-named drop-ins are set aside (and REF_SRC leads sys.path); afterwards the drop-ins are
    put back and the reference's modules live only in this shim's cache.

    NB: importing reseeds torch / numpy / random to 42 (reference generative_model_utils.py:22-26).
    """
    if not available():
        raise RuntimeError(f"reference sources not found at {REF_SRC}")
    if module_name in _loaded:
        return _loaded[module_name]
    for name in _STUBS:
        if name not in sys.modules:
            m = MagicMock(name=name)
            m.__path__ = []
            m.__all__ = []
            sys.modules[name] = m
    ref_names = {f[:-3] for f in os.listdir(REF_SRC) if f.endswith(".py")}
    aside = {n: sys.modules.pop(n) for n in list(sys.modules)
             if n in ref_names and not _from_reference(sys.modules[n])}
    sys.modules.update(_ref_modules)
    sys.path.insert(0, REF_SRC)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module(module_name)
    finally:
        sys.path.remove(REF_SRC)
        for n in list(sys.modules):
            if n in ref_names and _from_reference(sys.modules[n]):
                _ref_modules[n] = sys.modules.pop(n)
        sys.modules.update(aside)
    assert _from_reference(mod), f"{module_name} resolved to {getattr(mod, '__file__', None)}, not to the reference"
    _loaded[module_name] = mod
    return mod


def make_trainer(variant: str, n_genes: int, *, optimizer: str = "adam", hidden: int = 256,
                 latent: int = 256, embed: int = 256, seed: int = 0, dropout: float | None = 0.0,
                 negative_slope: float = 0.0, text_dim: int = 768, patch_dim: int = 1024, **kw):
    """Builds the reference trainer of `variant` on CPU exactly as its __main__ does (minus fit()).

    dropout=None keeps the reference's p=0.1; a number overrides every dropout probability in both
    nets (the only way to get run-to-run identical outputs).
    """
    import torch

    out_dir = tempfile.mkdtemp(prefix="gemmgan_ref_")
    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "vanilla":
            ref = load("vanilla_gan_unconditional")
            torch.manual_seed(seed)
            t = ref.WGAN_GP_nocond(input_dims=n_genes, latent_dims=latent, vocab_sizes=[],
                                   generator_dims=[hidden, hidden, n_genes],
                                   discriminator_dims=[hidden, hidden, 1], optimizer=optimizer,
                                   negative_slope=negative_slope, results_dire=out_dir, **kw)
            t.build_WGAN_GP_nocond()
        elif variant == "label":
            ref = load("benchmark_generative_model")
            torch.manual_seed(seed)
            t = ref.WGAN_GP_benchmark(input_dims=n_genes, latent_dims=latent, vocab_sizes=[10, 10],
                                      generator_dims=[hidden, hidden, n_genes],
                                      discriminator_dims=[hidden, hidden, 1], optimizer=optimizer,
                                      negative_slope=negative_slope, results_dire=out_dir, **kw)
            t.build_WGAN_GP()
        elif variant in ("concat", "concat_image"):
            ref = load("conditional_gan_concat")
            torch.manual_seed(seed)
            image = variant == "concat_image"
            t = ref.WGAN_GP(input_dims=n_genes, latent_dims=latent, embedding_dims=embed,
                            generator_dims=[hidden, hidden, n_genes], discriminator_dims=[hidden, hidden, 1],
                            input_embedding_dims=patch_dim if image else text_dim,
                            condition_on="image" if image else "text", optimizer=optimizer,
                            negative_slope=negative_slope, results_dire=out_dir, **kw)
            t.build_WGAN_GP()
        else:
            mod = {"paper": "conditional_gan_cross_attention_with_film",
                   "film": "conditional_gan_film", "cross": "conditional_gan_cross_attention",
                   "img": "conditional_gan_img_transformer", "attn": "conditional_gan_attention"}[variant]
            ref = load(mod)
            torch.manual_seed(seed)
            t = ref.WGAN_GP(input_dims=n_genes, latent_dims=latent, embedding_dims=embed,
                            generator_dims=[hidden, hidden, n_genes],
                            discriminator_dims=[hidden, hidden, 1], optimizer=optimizer,
                            negative_slope=negative_slope, results_dire=out_dir,
                            text_embedding_dims=text_dim, patches_embedding_dims=patch_dim, **kw)
            t.build_WGAN_GP()
        t.init_train()
    if dropout is not None:
        set_dropout(t.gen, dropout)
        set_dropout(t.disc, dropout)
    return t


def set_dropout(module, p: float) -> None:
    import torch

    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = p
