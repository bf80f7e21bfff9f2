"""Free-standing use of the drop-in modules: `generator(z, ...)` / `discriminator(x, ...)` outside a WGAN_GP trainer,
differentiable through torch.autograd (SURVEY.md §8 b "who calls it": the reference's nets are ordinary nn.Modules,
src/conditional_gan_cross_attention_with_film.py:97-233).

The forward is the engine's (gg_engine_generate_keep / gg_engine_critic_keep), the backward the engine's hand-written
one (gg_engine_generate_backward / gg_engine_critic_backward) wrapped in ONE torch.autograd.Function per call: the
module's parameters and the first input (z / gene profiles) receive gradients exactly as with a torch module —
`loss.backward()`, `torch.autograd.grad`, any torch optimizer stepping the parameters in place. First order only: a
double backward (the gradient penalty written with `create_graph=True`) raises; that computation is
WGAN_GP.gradient_penalty / train_disc of the drop-in trainers (closed-form double backward inside the engine).

An engine owns one (generator, critic) pair, so a free-standing net gets a never-used partner of matching sizes (built
under a forked RNG: the caller's random stream is untouched). Every forward that will be back-propagated holds its own
engine (its kept activations) until its graph is released, so `D(fake)`, `D(real)` and a joint `loss.backward()` work;
engines are pooled per input shape and reused.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .runtime import Engine, FlatNet


def _partner(module):
    """A net of the other role with sizes that fit `module`'s engine (weights never used)."""
    from . import models as M

    v, gen = module._variant, module._role == "gen"
    blocks = module.trunk_blocks()
    H = blocks[0][0].out_features
    slope = float(module.negative_slope)
    G = module.final_layer.out_features if gen else module.vector_dims
    L = module.latent_dims if gen else 16
    pairs = {"paper": (M.PaperGenerator, M.PaperDiscriminator), "cross": (M.CrossGenerator, M.CrossDiscriminator),
             "film": (M.FilmGenerator, M.FilmDiscriminator), "img": (M.ImgGenerator, M.ImgDiscriminator),
             "attn": (M.AttnGenerator, M.AttnDiscriminator), "concat": (M.ConcatGenerator, M.ConcatDiscriminator),
             "label": (M.LabelGenerator, M.LabelDiscriminator), "vanilla": (M.VanillaGenerator, M.VanillaDiscriminator)}
    GenC, DiscC = pairs[v]
    with torch.random.fork_rng(devices=[]):
        if v in ("paper", "cross", "film", "img", "attn"):
            E, Dt, Dp = module.embedding_dims, module.text_embedding_dims, module.patches_embedding_dims
            other = DiscC(G, E, [H, H, 1], Dt, Dp, slope) if gen else GenC(L, E, [H, H, G], Dt, Dp, slope)
        elif v == "concat":
            E, Di, ct = module.embedding_dims, module.input_embedding_dims, module.condition_type
            other = DiscC(G, Di, E, [H, H, 1], ct, slope) if gen else GenC(L, Di, E, [H, H, G], ct, slope)
        else:
            vs = list(module.vocab_sizes)
            other = DiscC(G, [], vs, [H, H, 1], slope) if gen else GenC(L, [], vs, [H, H, G], slope)
    return other, G, L, H


class _Lease:
    """Keeps one engine (the activations its last *_keep forward stored) out of the pool while a graph may still
    back-propagate through it; returns it when the graph's context is released."""

    def __init__(self, pool: List[Engine], eng: Engine):
        self.pool, self.eng = pool, eng

    def __del__(self):
        try:
            self.pool.append(self.eng)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class _ModuleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, eng, lease, x, training, *params):
        gen = owner.role == "gen"
        out = eng.generate_keep(x, training) if gen else eng.critic_keep(x, training)
        ctx.owner, ctx.eng, ctx.lease = owner, eng, lease
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        owner, eng = ctx.owner, ctx.eng
        want_dx = ctx.needs_input_grad[3]
        dout = dout.contiguous()
        if owner.role == "gen":
            dx = eng.generate_backward(dout, want_dx)
        else:
            dx = eng.critic_backward(dout, want_dx)
        flat = owner.flat
        grads = []
        for i, (slot, p) in enumerate(owner.param_list):
            if ctx.needs_input_grad[5 + i]:
                o = flat.offsets[slot]
                grads.append(flat.grads[o:o + p.numel()].view(p.shape).clone())
            else:
                grads.append(None)
        return (None, None, None, dx, None) + tuple(grads)


class Standalone:
    """Owner object of a free-standing module (module._gg_owner): flat parameter buffers, engines per input shape."""

    def __init__(self, module):
        params = list(module.parameters())
        dev = params[0].device
        _lib.require_cuda_tensor_device(dev, f"{type(module).__name__}.forward")
        self.module, self.role, self.variant, self.device = module, module._role, module._variant, dev
        self.partner, self.G, self.L, self.H = _partner(module)
        self.partner.to(dev)
        for p in self.partner.parameters():
            p.requires_grad_(False)
        self.flat = FlatNet(module, dev, "rms_prop", bind_grads=False)
        self.flat_partner = FlatNet(self.partner, dev, "rms_prop", bind_grads=False)
        self.param_list: List[Tuple[int, torch.nn.Parameter]] = sorted(self.flat.slots.items())
        self.pools: Dict[tuple, List[Engine]] = {}
        v = self.variant
        towers = v in ("paper", "cross", "film", "img")
        self.dropout_p = float(module.patches_transformer_layer.dropout.p) if towers else 0.0
        self.seed = int(getattr(module, "dropout_seed", 0))
        self.n_engines = 0

    # ---------------------------------------------------------------------------------- engines
    def moved(self) -> bool:
        """True when the parameters no longer live in the flat buffer (module.to(...), p.data reassigned)."""
        base = self.flat.params.data_ptr()
        return any(p.data_ptr() != base + 4 * self.flat.offsets[s] for s, p in self.param_list)

    def _shape(self, cond) -> dict:
        m, v = self.module, self.variant
        if v == "vanilla":
            return dict(E=0, H=self.H, Dt=0, Dp=0, P=0, T=0)
        if v == "label":
            return dict(E=m.categorical_embedded_dims, H=self.H, Dt=m.vocab_sizes[0], Dp=m.vocab_sizes[1], P=1, T=1)
        if v == "concat":
            return dict(E=m.embedding_dims, H=self.H, Dt=m.input_embedding_dims, Dp=m.input_embedding_dims, P=1, T=1,
                        tower_bias=True)
        if v in ("paper", "cross"):
            patches, _, text, _ = cond
            P, T = patches.shape[1], text.shape[1]
        else:
            _, patches, _ = cond
            P, T = patches.shape[1], 1
        return dict(E=m.embedding_dims, H=self.H, Dt=m.text_embedding_dims, Dp=m.patches_embedding_dims, P=P, T=T,
                    tower_bias=v in ("paper", "attn"))

    def _take(self, B: int, cond) -> Tuple[Engine, List[Engine]]:
        s = self._shape(cond)
        key = (B, s["P"], s["T"])
        pool = self.pools.setdefault(key, [])
        if pool:
            return pool.pop(), pool
        gen, disc = (self.flat, self.flat_partner) if self.role == "gen" else (self.flat_partner, self.flat)
        # every engine carries its own Philox (seed, step) state: engines in flight at the same time (D(fake) and D(real)
        # of one loss) must not replay each other's dropout masks
        seed = (self.seed + 0x9E3779B97F4A7C15 * self.n_engines) % (1 << 64)
        self.n_engines += 1
        eng = Engine(variant=self.variant, B=B, G=self.G, L=self.L, gen=gen, disc=disc,
                     slope=float(self.module.negative_slope), dropout_p=self.dropout_p, gp_weight=10.0, clip_d=0.0,
                     clip_g=0.0, optimizer="rms_prop", seed=seed, device=self.device, **s)
        return eng, pool

    def _stage(self, eng: Engine, cond) -> None:
        v = self.variant
        if v == "vanilla":
            return
        if v == "label":
            eng.set_labels(cond[0], cond[1])
            return
        dev = self.device
        if v in ("paper", "cross"):
            patches, ppad, text, tpad = cond
            eng.set_batch(patches=patches.to(dev), patch_pad=ppad.to(dev), text=text.to(dev),
                          text_pad=None if tpad is None else tpad.to(dev))
            return
        text, patches, ppad = cond
        if v == "concat":
            vec = text.to(dev) if self.module.condition_type == "text" else \
                eng.masked_mean_rows(patches.to(dev), ppad.to(dev))
            eng.set_batch(text=vec)
            return
        eng.set_batch(patches=patches.to(dev), patch_pad=ppad.to(dev), text=text.to(dev), text_pad=None)

    # ---------------------------------------------------------------------------------- forward
    def _module_forward(self, module, x, *cond):
        assert module is self.module
        x = x.to(self.device)
        eng, pool = self._take(x.shape[0], cond)
        eng.sync_params()                      # optimizer.step() / load_state_dict since the last call
        self._stage(eng, cond)
        training = bool(module.training)
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for _, p in self.param_list))
        if self.variant == "attn" and self.role == "gen" and training:
            module.attn_bn.num_batches_tracked += 1
        if not needs_graph:
            out = eng.generate(x, training=training) if self.role == "gen" else eng.critic(x, training=training)
            pool.append(eng)
            return out
        return _ModuleFn.apply(self, eng, _Lease(pool, eng), x, training, *[p for _, p in self.param_list])


def owner_of(module) -> Standalone:
    """The module's free-standing owner, (re)built when the parameters were moved since the last call."""
    own: Optional[Standalone] = getattr(module, "_gg_standalone", None)
    if own is None or own.moved():
        own = Standalone(module)
        object.__setattr__(module, "_gg_standalone", own)
    return own
