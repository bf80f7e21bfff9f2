"""Host-side runtime: flat parameter buffers and the ctypes wrapper of gg_engine.

PyTorch provides device memory and streams; every computation is a C-ABI call into
libgemmgan_sm100a.so. Nothing here has a CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
from torch import nn

from . import _abi_decl as A
from . import _lib

_ALIGN = 64  # elements (256 B): every tensor starts 16-byte aligned for 128-bit accesses


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(None if t is None else t.data_ptr())


class FlatNet:
    """Moves a module's parameters into one flat fp32 buffer (nn.Parameters become views), with a
    matching flat gradient buffer and optimizer-state buffers. Tensors the forward never uses
    (the prototype encoder layer the reference registers at :114) are placed behind `n_used`:
    they get no gradient (`grad is None`, as in the reference) and are never updated."""

    def __init__(self, module: nn.Module, device: torch.device, optimizer: str, bind_grads: bool = True):
        # bind_grads=False (free-standing modules driven by autograd): p.grad stays autograd's; the engine's gradient
        # buffer is only the place the backward kernels write to
        self.module = module
        self.bind_grads = bind_grads
        slots: Dict[int, nn.Parameter] = module.slot_table()
        used_ids = {id(p) for p in slots.values()}
        unused = [p for p in module.parameters() if id(p) not in used_ids]
        off, self.offsets = 0, {}
        for slot in sorted(slots):
            self.offsets[slot] = off
            off += (slots[slot].numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n_used = off
        tail = {}
        for p in unused:
            tail[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.params = torch.zeros(off, device=device, dtype=torch.float32)
        self.grads = torch.zeros(self.n_used, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n_used, device=device, dtype=torch.float32)
        self.exp_avg = (torch.zeros(self.n_used, device=device, dtype=torch.float32)
                        if optimizer in ("adam", "adamw") else None)
        self.step_count = torch.zeros(4, device=device, dtype=torch.float32)
        self.slots = slots
        with torch.no_grad():
            for slot, p in slots.items():
                o = self.offsets[slot]
                view = self.params[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = self.grads[o:o + p.numel()].view(p.shape) if bind_grads else None
            for p in unused:
                o = tail[id(p)]
                view = self.params[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = None
        # buffers (none in these models) and anything else follow the module to the device
        self._versions = self._version_sum()
        # Generation of the fp32 master weights. Every engine keeps its own bf16 shadows of them
        # (engine.cu::layout_shadows); an optimizer step run by ONE engine, or a write through PyTorch,
        # bumps this counter, and every engine refreshes its shadows when the generation it last saw
        # lags (Engine.sync_params) — engines of other batch sizes (last partial batch of an epoch, a
        # validation loader with its own batch size) must never run on stale matrices.
        self.generation = 0

    def _version_sum(self) -> int:
        return sum(p._version for p in self.slots.values())

    def poll_external_writes(self) -> int:
        """Bumps `generation` when someone wrote to the parameters through PyTorch (load_state_dict,
        manual edits, a broadcast) since the last poll; returns the current generation."""
        v = self._version_sum()
        if v != self._versions:
            self._versions = v
            self.generation += 1
        return self.generation

    def bump(self) -> int:
        """An engine's optimizer kernel has rewritten the master weights (the C kernels do not touch
        torch's version counters)."""
        self.generation += 1
        return self.generation

    def attach_optimizer(self, opt: "torch.optim.Optimizer") -> None:
        """Makes `opt.state_dict()` / `opt.load_state_dict()` meaningful although `opt.step()` is never called: the
        per-parameter state entries torch's RMSprop / Adam / AdamW keep (`square_avg` / `exp_avg`, `exp_avg_sq`,
        `step`; torch/optim/rmsprop.py, adam.py) become VIEWS of the flat state buffers the optimizer kernel
        updates, so torch.save(optimizer.state_dict()) captures the live state; load_state_dict copies the loaded
        tensors back into the flat buffers (torch replaces the state tensors on load) and re-attaches the views.
        Tensors the forward never uses (grad is None in the reference too) have no state, as in torch."""
        adam = self.exp_avg is not None

        def attach():
            opt.state.clear()
            for slot, p in self.slots.items():
                o, n = self.offsets[slot], p.numel()
                st = {"step": self.step_count[0:1].view(())}
                if adam:
                    st["exp_avg"] = self.exp_avg[o:o + n].view(p.shape)
                    st["exp_avg_sq"] = self.exp_avg_sq[o:o + n].view(p.shape)
                else:
                    st["square_avg"] = self.exp_avg_sq[o:o + n].view(p.shape)
                opt.state[p] = st

        def after_load(optimizer):
            with torch.no_grad():
                for slot, p in self.slots.items():
                    st = optimizer.state.get(p)
                    if not st:
                        continue
                    o, n = self.offsets[slot], p.numel()
                    if adam:
                        self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                        self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    else:
                        self.exp_avg_sq[o:o + n].copy_(st["square_avg"].reshape(-1))
                    self.step_count[0:1].copy_(torch.as_tensor(st["step"], dtype=torch.float32).reshape(1))
            attach()

        attach()
        opt.register_load_state_dict_post_hook(after_load)

    def reattach_grads(self) -> None:
        """optimizer.zero_grad(set_to_none=True) drops p.grad; point it back at the flat buffer."""
        for slot, p in self.slots.items():
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * self.offsets[slot]:
                o = self.offsets[slot]
                p.grad = self.grads[o:o + p.numel()].view(p.shape)

    def c_struct(self) -> A.NetBuffers:
        nb = A.NetBuffers()
        nb.params = self.params.data_ptr()
        nb.grads = self.grads.data_ptr()
        nb.exp_avg = None if self.exp_avg is None else self.exp_avg.data_ptr()
        nb.exp_avg_sq = self.exp_avg_sq.data_ptr()
        nb.step_count = self.step_count.data_ptr()
        nb.n_used = self.n_used
        for s in range(A.NSLOTS):
            nb.off[s] = self.offsets.get(s, -1)
        return nb


_OPT_IDS = {"rms_prop": A.OPT_RMSPROP, "adam": A.OPT_ADAM, "adamw": A.OPT_ADAMW}


class Engine:
    """One gg_engine: a (generator, critic) pair at a fixed per-rank batch size."""

    def __init__(self, *, variant: str, B: int, G: int, L: int, E: int, H: int, Dt: int, Dp: int, P: int, T: int,
                 gen: FlatNet, disc: FlatNet, slope: float, dropout_p: float, gp_weight: float, clip_d: float,
                 clip_g: float, optimizer: str, seed: int = 0, gemm_impl: int = _lib.IMPL_TCGEN05,
                 tower_bias: bool = True, device: Optional[torch.device] = None):
        from .models import VARIANT_IDS

        self.lib = _lib.lib()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        _lib.require_device(self.device.index or 0)
        cfg = A.ModelCfg()
        cfg.variant = VARIANT_IDS[variant]
        cfg.B, cfg.G, cfg.L, cfg.E, cfg.H = B, G, L, E, H
        cfg.Dt, cfg.Dp, cfg.P, cfg.T = Dt, Dp, P, T
        cfg.n_layers, cfg.n_heads, cfg.ffn = 2, 4, 2 * E
        cfg.tower_bias = int(tower_bias)
        cfg.slope, cfg.dropout_p, cfg.gp_weight = slope, dropout_p, gp_weight
        cfg.clip_d, cfg.clip_g, cfg.ln_eps = clip_d, clip_g, 1e-5
        cfg.optimizer = _OPT_IDS[optimizer]
        cfg.gemm_impl = gemm_impl
        cfg.seed = seed
        self.cfg = cfg
        self.variant, self.B, self.G, self.L = variant, B, G, L
        self.gen, self.disc = gen, disc
        nbytes = C.c_int64(0)
        _lib.check(self.lib.gg_engine_workspace_bytes(C.byref(cfg), C.byref(nbytes)))
        # the engine wants a 256-byte aligned workspace: CUDA allocations are (offset 0), host allocations of the
        # emulated test build are not
        raw = torch.empty(nbytes.value + 256, device=self.device, dtype=torch.uint8)
        off = (-raw.data_ptr()) % 256
        self.workspace = raw[off:off + nbytes.value]
        self._gen_c, self._disc_c = gen.c_struct(), disc.c_struct()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gg_engine_create(C.byref(cfg), C.byref(self._gen_c), C.byref(self._disc_c),
                                                 _ptr(self.workspace), nbytes.value, _stream(), C.byref(h)))
        self.handle = h
        if variant == "attn":
            # the generator's BatchNorm1d buffers stay torch's (state_dict / checkpoints); the kernel updates them in place
            bn = gen.module.attn_bn
            rm, rv = bn.running_mean, bn.running_var
            assert rm.device == self.device and rm.dtype == torch.float32 and rm.is_contiguous() and rv.is_contiguous()
            assert bn.momentum is not None and bn.track_running_stats and bn.affine
            _lib.check(self.lib.gg_engine_set_batchnorm(self.handle, _ptr(rm), _ptr(rv), float(bn.momentum),
                                                        float(bn.eps)))
            self._keep_bn = (rm, rv)
        sp = self.lib.gg_engine_stats(self.handle)
        self.stats = self._view(sp, A.STATS_COUNT, torch.float32)
        self._keep = None
        # fixed-address noise inputs so that a captured CUDA graph of the step can be replayed
        self.z_in = torch.empty(B, L, device=self.device, dtype=torch.float32)
        self.alpha_in = torch.empty(B, 1, device=self.device, dtype=torch.float32)
        self.z_all = self.alpha_all = None
        self.graphs, self.warmed = {}, set()
        # generation of each net's master weights this engine's bf16 shadows were made from (create refreshes both)
        self.seen = {A.NET_GEN: gen.poll_external_writes(), A.NET_DISC: disc.poll_external_writes()}

    def ensure_noise(self, n_critic: int) -> None:
        """Per-step noise buffers of one train() call: z_all [n_critic + 1, B, L], alpha_all [n_critic, B, 1]."""
        if self.z_all is None or self.z_all.shape[0] != n_critic + 1:
            self.z_all = torch.empty(n_critic + 1, self.B, self.L, device=self.device, dtype=torch.float32)
            self.alpha_all = torch.empty(n_critic, self.B, 1, device=self.device, dtype=torch.float32)
            self.graphs.clear()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.gg_engine_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _view(self, ptr: int, n: int, dtype: torch.dtype) -> torch.Tensor:
        off = ptr - self.workspace.data_ptr()
        size = n * torch.empty((), dtype=dtype).element_size()
        assert 0 <= off and off + size <= self.workspace.numel(), "pointer outside the engine workspace"
        return self.workspace[off:off + size].view(dtype)

    def buffer(self, name: str) -> torch.Tensor:
        """Named internal device buffer as a [rows, cols] tensor view (tests / diagnostics)."""
        rows, cols, ld, f32 = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        p = self.lib.gg_engine_buffer(self.handle, name.encode(), C.byref(rows), C.byref(cols), C.byref(ld),
                                      C.byref(f32))
        if not p:
            raise KeyError(name)
        dt = torch.float32 if f32.value else torch.bfloat16
        flat = self._view(p, (rows.value - 1) * ld.value + cols.value, dt)
        return torch.as_strided(flat, (rows.value, cols.value), (ld.value, 1))

    # ---- thin wrappers ---------------------------------------------------------------------
    def set_lanes(self, enabled: bool) -> None:
        """Side-stream concurrency inside the entry points (see gg_engine_set_lanes); drops captured graphs."""
        _lib.check(self.lib.gg_engine_set_lanes(self.handle, int(enabled)))
        self.graphs.clear()

    def refresh_shadows(self, net: Optional[int] = None) -> None:
        for n in ((A.NET_GEN, A.NET_DISC) if net is None else (net,)):
            _lib.check(self.lib.gg_engine_refresh_shadows(self.handle, n, _stream()))

    def sync_params(self) -> None:
        """Refreshes the bf16 shadows of every net whose master weights changed since this engine last saw
        them: by another engine's optimizer step, or by a write through PyTorch."""
        for net, flat in ((A.NET_GEN, self.gen), (A.NET_DISC, self.disc)):
            g = flat.poll_external_writes()
            if self.seen[net] != g:
                self.refresh_shadows(net)
                self.seen[net] = g

    sync_external_param_writes = sync_params

    def stepped(self, net: int) -> None:
        """This engine's optimizer kernel has updated `net` (and refreshed ITS shadows): the other engines lag."""
        flat = self.gen if net == A.NET_GEN else self.disc
        self.seen[net] = flat.bump()

    @staticmethod
    def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()

    @staticmethod
    def _u8(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = t.contiguous()
        return t.view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8)

    def _expect(self, what: str, t: Optional[torch.Tensor], *shape: int) -> None:
        """The library reads raw pointers with the engine's static sizes: a tensor of another size must fail here, with
        the sizes in the message (the reference's torch modules raise a shape error at the first matmul), not read out
        of bounds on the device."""
        if t is None:
            return
        n = 1
        for d in shape:
            n *= d
        if t.numel() != n or t.shape[0] != shape[0]:
            raise ValueError(f"{what}: expected shape {tuple(shape)} for this engine "
                             f"(variant {self.variant!r}, batch {self.B}), got {tuple(t.shape)}")
        d, e = t.device, self.device
        if d.type != e.type or (d.index is not None and e.index is not None and d.index != e.index):
            raise ValueError(f"{what} is on {d}, the engine on {e}")

    def set_batch(self, genes=None, patches=None, patch_pad=None, text=None, text_pad=None) -> None:
        c = self.cfg
        self._expect("gene expression", genes, self.B, self.G)
        self._expect("patches", patches, self.B, c.P, c.Dp)
        self._expect("patch padding mask", patch_pad, self.B, c.P)
        self._expect("text embedding", text, self.B, c.T, c.Dt)
        self._expect("text padding mask", text_pad, self.B, c.T)
        g, p, t = self._f32(genes), self._f32(patches), self._f32(text)
        pm, tm = self._u8(patch_pad), self._u8(text_pad)
        self._keep = (g, p, t, pm, tm)  # keep alive until the enqueued casts have run
        _lib.check(self.lib.gg_engine_set_batch(self.handle, _ptr(g), _ptr(p), _ptr(pm), _ptr(t), _ptr(tm), _stream()))

    def set_labels(self, labels0: torch.Tensor, labels1: torch.Tensor) -> None:
        """Label-conditioned baseline: the two categorical covariates of the batch, int64 [B] on the device."""
        y0 = labels0.reshape(-1).to(device=self.device, dtype=torch.int64).contiguous()
        y1 = labels1.reshape(-1).to(device=self.device, dtype=torch.int64).contiguous()
        assert y0.numel() == self.B and y1.numel() == self.B
        self._keep_labels = (y0, y1)
        _lib.check(self.lib.gg_engine_set_labels(self.handle, _ptr(y0), _ptr(y1), _stream()))

    def disc_grads(self, z: torch.Tensor, alpha: torch.Tensor, training: bool = True, phase: int = 0,
                   gen_eval: bool = False) -> None:
        """phase 0 = whole step; 1 = forward + trunk backward; 2 = tower backward (gg_engine_disc_grads_phase).
        gen_eval: the generator forward inside the step runs in eval mode (GG_TRAIN_GEN_EVAL)."""
        z, alpha = self._f32(z), self._f32(alpha)
        assert z.shape == (self.B, self.L) and alpha.numel() == self.B
        mode = int(bool(training)) | (A.TRAIN_GEN_EVAL if gen_eval else 0)
        if phase == 0:
            _lib.check(self.lib.gg_engine_disc_grads(self.handle, _ptr(z), _ptr(alpha), mode, _stream()))
        else:
            _lib.check(self.lib.gg_engine_disc_grads_phase(self.handle, _ptr(z), _ptr(alpha), mode, phase, _stream()))

    def gen_grads(self, z: torch.Tensor, training: bool = True, phase: int = 0) -> None:
        z = self._f32(z)
        assert z.shape == (self.B, self.L)
        if phase == 0:
            _lib.check(self.lib.gg_engine_gen_grads(self.handle, _ptr(z), int(training), _stream()))
        else:
            _lib.check(self.lib.gg_engine_gen_grads_phase(self.handle, _ptr(z), int(training), phase, _stream()))

    def lanes_signal(self, stream: torch.cuda.Stream) -> None:
        """`stream` waits for the engine's side lanes (gg_engine_lanes_signal)."""
        _lib.check(self.lib.gg_engine_lanes_signal(self.handle, C.c_void_p(stream.cuda_stream)))

    def optim_step(self, net: int, lr: float) -> None:
        _lib.check(self.lib.gg_engine_optim_step(self.handle, net, float(lr), _stream()))

    def generate(self, z: torch.Tensor, training: bool = False) -> torch.Tensor:
        self._expect("z", z, self.B, self.L)
        z = self._f32(z)
        out = torch.empty(self.B, self.G, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_engine_generate(self.handle, _ptr(z), _ptr(out), int(training), _stream()))
        return out

    # ---- module-level autograd (gemmgan_b200/standalone.py): forward keeping the backward's tensors + first-order backward
    def generate_keep(self, z: torch.Tensor, training: bool) -> torch.Tensor:
        self._expect("z", z, self.B, self.L)
        z = self._f32(z)
        out = torch.empty(self.B, self.G, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_engine_generate_keep(self.handle, _ptr(z), _ptr(out), int(training), _stream()))
        return out

    def generate_backward(self, dout: torch.Tensor, want_dz: bool) -> Optional[torch.Tensor]:
        d = self._f32(dout)
        dz = torch.empty(self.B, self.L, device=self.device, dtype=torch.float32) if want_dz else None
        _lib.check(self.lib.gg_engine_generate_backward(self.handle, _ptr(d), _ptr(dz), _stream()))
        return dz

    def critic_keep(self, genes: torch.Tensor, training: bool) -> torch.Tensor:
        self._expect("gene expression", genes, self.B, self.G)
        g = self._f32(genes)
        out = torch.empty(self.B, 1, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_engine_critic_keep(self.handle, _ptr(g), _ptr(out), int(training), _stream()))
        return out

    def critic_backward(self, dscore: torch.Tensor, want_dx: bool) -> Optional[torch.Tensor]:
        d = self._f32(dscore.reshape(-1))
        dx = torch.empty(self.B, self.G, device=self.device, dtype=torch.float32) if want_dx else None
        _lib.check(self.lib.gg_engine_critic_backward(self.handle, _ptr(d), _ptr(dx), _stream()))
        return dx

    def gradient_penalty(self, real, fake, alpha, training: bool = True) -> torch.Tensor:
        self._expect("real data", real, self.B, self.G)
        self._expect("fake data", fake, self.B, self.G)
        self._expect("alpha", alpha, self.B, 1)
        r, f, a = self._f32(real), self._f32(fake), self._f32(alpha)
        out = torch.empty((), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_engine_gradient_penalty(self.handle, _ptr(r), _ptr(f), _ptr(a), int(training),
                                                       _ptr(out), _stream()))
        return out

    def masked_mean_rows(self, x: torch.Tensor, pad: Optional[torch.Tensor]) -> torch.Tensor:
        """[B, P, D] fp32 -> [B, D]: mean over the rows whose pad flag is False (gg_masked_mean_rows)."""
        x, pm = self._f32(x), self._u8(pad)
        B, P, D = x.shape
        out = torch.empty(B, D, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_masked_mean_rows(_ptr(x), _ptr(pm), _ptr(out), B, P, D, _stream()))
        return out

    def gp_step(self, real, fake, alpha, out: Optional[torch.Tensor] = None) -> None:
        """GP value + gp_weight * dGP/d{W1, W2, w3} into the critic's gradient buffer (gg_engine_gp_step)."""
        r, f, a = self._f32(real), self._f32(fake), self._f32(alpha)
        _lib.check(self.lib.gg_engine_gp_step(self.handle, _ptr(r), _ptr(f), _ptr(a), _ptr(out), _stream()))

    def critic(self, genes: torch.Tensor, training: bool = False) -> torch.Tensor:
        self._expect("gene expression", genes, self.B, self.G)
        g = self._f32(genes)
        out = torch.empty(self.B, 1, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.gg_engine_critic(self.handle, _ptr(g), _ptr(out), int(training), _stream()))
        return out
