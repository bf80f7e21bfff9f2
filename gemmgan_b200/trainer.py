"""Trainer behind the drop-in WGAN_GP classes: same constructor kwargs, methods and attributes as
the reference trainers, with train_disc / train_gen executed by the sm_100a engine.

Reference: class WGAN_GP, src/conditional_gan_cross_attention_with_film.py:256-477 (paper model),
src/conditional_gan_film.py:237-445 (film), class WGAN_GP_nocond src/vanilla_gan_unconditional.py:211-431.

Differences that are deliberate (documented in DESIGN.md / INTEGRATION.md):
  * the per-step `.item()` host syncs (:421-423, :461) are replaced by an async copy of the engine's
    stats vector into pinned memory; `d_batch_loss`, `g_batch_loss`, `disc_loss`, `gen_loss` are
    properties that synchronise on first read;
  * `optimizer_disc` / `optimizer_gen` are real torch.optim objects over the same Parameters (so
    `param_groups[i]['lr']` scheduling in fit() works) but their `.step()` is never called: the
    update is the engine's fused flat-buffer kernel;
  * data parallelism (absent in the reference): when torch.distributed is initialised the flat
    gradient buffer is all-reduced (mean) between the backward and the optimizer kernel.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import _abi_decl as A
from . import _lib
from .runtime import Engine, FlatNet


def wasserstein_loss(y_pred, y_true):
    return torch.mean(y_pred * y_true)


def G_loss(fake_labels):
    return wasserstein_loss(fake_labels, -torch.ones_like(fake_labels))


def D_loss(real_labels, fake_labels):
    loss_real = wasserstein_loss(-torch.ones_like(real_labels), real_labels)
    loss_fake = wasserstein_loss(torch.ones_like(fake_labels), fake_labels)
    return loss_real + loss_fake, loss_real, loss_fake


def save_numpy(file, data):
    """The scripts' helper for the .npy dumps (src/conditional_gan_cross_attention_with_film.py:28-30)."""
    with open(file, 'wb') as f:
        np.save(f, data)


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class TrainerBase:
    """Variant-independent machinery. Subclasses define `variant`, build the nets and translate the
    reference's argument orders into the engine's (genes, patches, patch_pad, text, text_pad)."""

    variant = "paper"
    clip_d: Optional[float] = None
    clip_g: Optional[float] = None

    # ---- construction -------------------------------------------------------------------
    def _init_common(self, input_dims, latent_dims, generator_dims, discriminator_dims, negative_slope, is_bn,
                     lr_d, lr_g, optimizer, gp_weight, p_aug, norm_scale, train, n_critic, freq_print,
                     freq_compute_test, freq_visualize_test, patience, normalization, log2, rpm, results_dire):
        self.input_dims = input_dims
        self.latent_dims = latent_dims
        self.generator_dims = generator_dims
        self.discriminator_dims = discriminator_dims
        self.negative_slope = negative_slope
        self.is_bn = is_bn
        self.gp_weight = gp_weight
        self.isTrain = train
        self.p_aug = p_aug
        self.norm_scale = norm_scale
        self.n_genes = input_dims
        self.n_critic = n_critic
        self.freq_print = freq_print
        self.freq_compute_test = freq_compute_test
        self.freq_visualize_test = freq_visualize_test
        self.result_dire = self.results_dire = results_dire
        if results_dire:
            os.makedirs(results_dire, exist_ok=True)
            self.results_dire_fig = os.path.join(results_dire, "figures")
            os.makedirs(self.results_dire_fig, exist_ok=True)
        self.dend = False
        self.lr_d, self.lr_g = lr_d, lr_g
        self.optimizer = optimizer
        self.patience = patience
        if p_aug != 0:
            # the reference's p_aug branch reads an undefined variable (:401) and cannot run
            raise NotImplementedError("p_aug != 0 is dead code in the reference (NameError at :401)")
        if not torch.cuda.is_available():
            raise RuntimeError("gemmgan_b200 needs a CUDA device of compute capability 10.x (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.loss_dict = {"d loss": [], "d real loss": [], "d fake loss": [], "g loss": []}
        self.corr_scores, self.corr_dend_scores = {}, {}
        self.precision_scores, self.recall_scores = {}, {}
        self.normalization, self.log2, self.rpm = normalization, log2, rpm
        self.dropout_p = 0.1          # nn.TransformerEncoderLayer(dropout=0.1) in the reference towers
        self.dropout_seed = 0
        self.gemm_impl = _lib.IMPL_TCGEN05
        self.dp_overlap = os.environ.get("GEMMGAN_DP_OVERLAP", "1") != "0"
        self.dp_graph_collectives = os.environ.get("GEMMGAN_DP_GRAPH", "1") != "0"
        self.dp_stage_buckets = os.environ.get("GEMMGAN_DP_STAGES", "1") != "0"
        self.dp_global_noise = False   # draw z/alpha for the GLOBAL batch and slice (N-rank == 1-rank parity)
        self.use_cuda_graphs = os.environ.get("GEMMGAN_CUDA_GRAPHS", "1") != "0"
        self.noise_upfront = os.environ.get("GEMMGAN_NOISE_UPFRONT", "1") != "0"
        # one graph per train() call instead of one per optimizer step: measured 7.31 vs 7.27 ms on the device and
        # 7.42 vs 7.51 ms end to end (N=1) -- within noise, so the per-step graphs (validated at N = 1..8) stay default
        self.whole_call_graph = os.environ.get("GEMMGAN_WHOLE_CALL_GRAPH", "0") != "0"
        self._in_capture = False
        self.unique_graphs = False
        self._graph_seq = 0
        self._replay_events = []
        self._copy_stream, self._prefetched, self._prefetch_event = None, {}, None
        self._engines = {}
        self._flat_gen = self._flat_disc = None
        self._pinned = None
        self._events = {}
        self._noise_gen = None

    def init_train(self):
        opt = self.optimizer.lower()
        if opt == "rms_prop":
            mk = lambda ps, lr: torch.optim.RMSprop(ps, lr=lr)
        elif opt == "adam":
            mk = lambda ps, lr: torch.optim.Adam(ps, lr=lr, betas=(0.9, 0.99))
        elif opt == "adamw":
            mk = lambda ps, lr: torch.optim.AdamW(ps, lr=lr, betas=(0.9, 0.99), weight_decay=0.01)
        else:
            raise ValueError(f"unknown optimizer {self.optimizer!r}")
        self.optimizer_disc = mk(self.disc.parameters(), self.lr_d)
        self.optimizer_gen = mk(self.gen.parameters(), self.lr_g)
        # the update itself is the engine's kernel; optimizer state lives in the flat buffers, which
        # optimizer.state_dict() / load_state_dict() see through views (FlatNet.attach_optimizer)
        self._flatten()
        self._flat_disc.attach_optimizer(self.optimizer_disc)
        self._flat_gen.attach_optimizer(self.optimizer_gen)

    def _attach(self, gen, disc):
        self.gen, self.disc = gen.to(self.device), disc.to(self.device)
        self.gen._gg_owner = self
        self.disc._gg_owner = self
        # new networks (fit() builds again, as the reference's does at :620-623): the flat buffers, engines and
        # captured graphs of the previous pair must not outlive it, or the engine would keep training the old modules
        self._flat_gen = self._flat_disc = None
        self._engines.clear()

    def _flatten(self):
        if self._flat_gen is None:
            opt = self.optimizer.lower()
            self._flat_gen = FlatNet(self.gen, self.device, opt)
            self._flat_disc = FlatNet(self.disc, self.device, opt)
            self._engines.clear()
            d = _dist()
            if d is not None:
                # replicas must start from the same weights whatever each rank's torch seed was: rank 0's go to all
                # (optimizer state starts at zero everywhere). The flat buffers are written in place, so the
                # nn.Parameter views follow; engines are created afterwards and build their shadows from them.
                for flat in (self._flat_gen, self._flat_disc):
                    d.broadcast(flat.params, src=0)
                    flat.bump()

    def _shape_cfg(self):
        raise NotImplementedError

    def _engine(self, B: int) -> Engine:
        """Engines are per batch size (fixed buffers); they all share the flat parameter buffers."""
        self._flatten()
        eng = self._engines.get(B)
        if eng is None:
            s = self._shape_cfg()
            eng = Engine(variant=self.variant, B=B, G=self.n_genes, L=self.latent_dims, gen=self._flat_gen,
                         disc=self._flat_disc, slope=float(self.negative_slope), dropout_p=float(self.dropout_p),
                         gp_weight=float(self.gp_weight), clip_d=float(self.clip_d or 0.0),
                         clip_g=float(self.clip_g or 0.0), optimizer=self.optimizer.lower(),
                         seed=int(self.dropout_seed), gemm_impl=self.gemm_impl, device=self.device, **s)
            self._engines[B] = eng
        else:
            eng.sync_params()
        return eng

    def _check_equal_rank_batches(self, eng: Engine) -> None:
        """Data parallel, once per TRAINING engine (first optimizer step, where every rank is present): the all-reduce
        averages per-rank gradient MEANS with equal weight, which is the global-batch gradient only when every rank
        holds the same number of rows (e.g. DistributedSampler(drop_last=True)). Not done at engine creation: rank 0
        alone creates engines while it evaluates (generate_samples_all over a validation loader), and a collective
        there would pair up with the other ranks' gradient all-reduces."""
        d = _dist()
        if d is None or getattr(eng, "_dp_batch_checked", False) or self._in_capture:
            return
        sizes = torch.zeros(d.get_world_size(), dtype=torch.int64, device=self.device)
        sizes[d.get_rank()] = eng.B
        d.all_reduce(sizes)
        if int(sizes.min()) != int(sizes.max()):
            raise ValueError(f"data-parallel ranks must use equal per-rank batch sizes, got {sizes.tolist()}")
        eng._dp_batch_checked = True

    # ---- host -> device staging ----------------------------------------------------------
    def prefetch(self, *tensors):
        """Starts the host->device copies of the NEXT batch's tensors on a copy stream, so that they overlap the
        step that is running; the next train() / train_disc() call that receives these same host tensors picks the
        device copies up (pass pinned tensors: pageable ones are copied synchronously by CUDA)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        self._prefetched.clear()
        with torch.cuda.stream(self._copy_stream):
            for t in tensors:
                if isinstance(t, torch.Tensor) and not t.is_cuda:
                    self._prefetched[(t.data_ptr(), tuple(t.shape), t.dtype)] = t.to(self.device, non_blocking=True)
            self._prefetch_event = torch.cuda.Event()
            self._prefetch_event.record()

    @staticmethod
    def _lookahead(loader):
        """(batch, next batch or None) pairs: fit() hands the next batch to train(prefetch=...)."""
        it = iter(loader)
        cur = next(it, None)
        while cur is not None:
            nxt = next(it, None)
            yield cur, nxt
            cur = nxt

    def _dev(self, t):
        """Device copy of one batch tensor: the prefetched copy when there is one, else an async copy now."""
        if t is None or t.is_cuda:
            return t
        hit = self._prefetched.pop((t.data_ptr(), tuple(t.shape), t.dtype), None)
        if hit is None:
            return t.to(self.device, non_blocking=True)
        cur = torch.cuda.current_stream()
        cur.wait_event(self._prefetch_event)
        hit.record_stream(cur)
        return hit

    # ---- noise --------------------------------------------------------------------------
    def _normal(self, B):
        d = _dist()
        if d is not None and self.dp_global_noise:
            w, r = d.get_world_size(), d.get_rank()
            z = torch.normal(0, 1, size=(w * B, self.latent_dims), device=self.device)
            return z[r * B:(r + 1) * B].contiguous()
        return torch.normal(0, 1, size=(B, self.latent_dims), device=self.device)

    def _alpha(self, B):
        d = _dist()
        if d is not None and self.dp_global_noise:
            w, r = d.get_world_size(), d.get_rank()
            a = torch.rand(w * B, 1, device=self.device)
            return a[r * B:(r + 1) * B].contiguous()
        return torch.rand(B, 1, device=self.device)

    # ---- stats readback -----------------------------------------------------------------
    def _snapshot(self, eng: Engine, key: str):
        if self._pinned is None:
            self._pinned = {k: torch.zeros(A.STATS_COUNT, dtype=torch.float32).pin_memory() for k in ("d", "g")}
        self._pinned[key].copy_(eng.stats, non_blocking=True)
        if not self._in_capture:   # (inside a whole-call capture the event is recorded after the replay)
            self._mark(key)

    def _mark(self, key: str):
        ev = self._events.get(key)
        if ev is None:
            ev = self._events[key] = torch.cuda.Event()
        ev.record()

    def _stats(self, key: str) -> np.ndarray:
        ev = self._events.get(key)
        if ev is None:
            raise AttributeError("no training step has run yet")
        ev.synchronize()
        return self._pinned[key].numpy()

    @property
    def d_batch_loss(self):
        s = self._stats("d")
        return np.array([float(s[A.STAT_LOSS_REAL] + s[A.STAT_LOSS_FAKE]), float(s[A.STAT_LOSS_REAL]),
                         float(s[A.STAT_LOSS_FAKE])])

    @property
    def g_batch_loss(self):
        return np.array([float(self._stats("g")[A.STAT_G_LOSS])])

    @property
    def disc_loss(self):
        return torch.tensor(float(self._stats("d")[A.STAT_D_LOSS]))

    @property
    def gen_loss(self):
        return torch.tensor(float(self._stats("g")[A.STAT_G_LOSS]))

    @property
    def last_gp(self):
        return float(self._stats("d")[A.STAT_GP])

    # ---- the hot path -------------------------------------------------------------------
    def _lr(self, opt):
        return opt.param_groups[0]["lr"]

    def _replay(self, eng: Engine, key, body):
        """Runs `body` (which only enqueues library kernels on the current stream). With
        use_cuda_graphs the kernel sequence is captured once per key and replayed afterwards: the
        step is ~200 short kernels, so launch latency would otherwise dominate.
        The first call per engine runs eagerly (one-time lazy initialisation inside the library)."""
        if not self.use_cuda_graphs or self._in_capture:
            body()          # eager mode, or already inside the capture of a whole train() call
            return
        if self.unique_graphs:  # diagnostics: one capture per step (bench.py's per-launch event timing)
            self._graph_seq += 1
            key = key + (self._graph_seq,)
        g = eng.graphs.get(key)
        if g is None:
            # 'd' / 'g' (+ 'e' = critic step with an eval-mode generator): the library's one-time lazy initialisation is
            # per step kind, and must not happen inside a capture
            kind = key[0][:1] + ("e" if "e" in key[0].split("_")[0][1:] else "")
            if kind not in eng.warmed:
                eng.warmed.add(kind)
                body()
                return
            g = torch.cuda.CUDAGraph()
            n0 = eng.lib.gg_launch_count(0)
            with torch.cuda.graph(g):
                body()
            n_kernels = eng.lib.gg_launch_count(0) - n0
            eng.lib.gg_launch_count_add(-n_kernels)     # capture enqueues nothing
            g = eng.graphs[key] = (g, n_kernels)
        if self.unique_graphs:  # diagnostics: device time of each replay (bench.py)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g[0].replay()
            e1.record()
            self._replay_events.append((e0, e1))
        else:
            g[0].replay()
        eng.lib.gg_launch_count_add(g[1])

    def _buckets(self, flat: FlatNet):
        """Trunk bucket (finished first by the hand-written backward) + tower bucket of one net's flat gradients."""
        from .ddp import GradBuckets, plan_buckets

        gb = getattr(flat, "_buckets", None)
        if gb is None:
            trunk = (A.P_TR0_W, A.P_TR0_B, A.P_TR1_W, A.P_TR1_B, A.P_FIN_W, A.P_FIN_B)
            gb = flat._buckets = GradBuckets(flat.grads, plan_buckets(flat.offsets, flat.n_used, trunk))
        return gb

    def _stage_buckets(self, flat: FlatNet):
        from .ddp import GradBuckets, plan_stage_buckets

        sb = getattr(flat, "_stage_buckets", None)
        if sb is None:
            trunk = (A.P_TR0_W, A.P_TR0_B, A.P_TR1_W, A.P_TR1_B, A.P_FIN_W, A.P_FIN_B)
            cross = tuple(range(A.P_P2T_IN_W, A.P_T2P_OUT_B + 1))
            plan = plan_stage_buckets(flat.offsets, flat.n_used, trunk, A.P_LAYER0, A.L_COUNT, 2, cross)
            sb = flat._stage_buckets = GradBuckets(flat.grads, [b for _, b in plan])
            sb.plan = plan
        return sb

    def _step(self, eng: Engine, tag, net, flat, lr, grads_fn):
        """grads_fn(phase): phase 0 = whole backward, 1 = forward + trunk backward, 2 = tower backward."""
        lr = float(lr)
        self._check_equal_rank_batches(eng)
        try:
            self._step_inner(eng, tag, net, flat, lr, grads_fn)
        finally:
            eng.stepped(net)    # master weights moved: engines of other batch sizes refresh their shadows

    def _step_inner(self, eng: Engine, tag, net, flat, lr, grads_fn):
        if _dist() is None:
            def body():
                grads_fn(0)
                eng.optim_step(net, lr)
            self._replay(eng, (tag, lr), body)
            return
        # data parallel: the trunk bucket (over half of the net: one [H, G] matrix) is all-reduced on the
        # communication stream while the fusion-tower backward runs; the tower bucket follows it; the
        # optimizer kernel waits for both
        if self.dp_overlap and self.dp_graph_collectives and self.dp_stage_buckets and self.variant != "vanilla":
            # staged backward: trunk, cross-attention tail, each encoder layer and the embedding tensors are reduced
            # as soon as their producers have been enqueued; the side lanes are not joined in between (the
            # communication stream waits for them directly), only the last, small bucket is exposed
            sb = self._stage_buckets(flat)
            stages = {st: i for i, (st, _) in enumerate(sb.plan)}
            n_stages = 2 + 2   # head, two encoder layers, embedding tail

            def body():
                nj = A.PHASE_NO_JOIN
                grads_fn(1 | nj)
                if -1 in stages:
                    sb.reduce(stages[-1], eng.lanes_signal)
                for st in range(n_stages):
                    last = st == n_stages - 1
                    grads_fn((A.PHASE_STAGE0 + st) | (0 if last else nj))
                    if st in stages:
                        sb.reduce(stages[st], None if last else eng.lanes_signal)
                sb.wait()
                eng.optim_step(net, lr)
            self._replay(eng, (tag + "_dps", lr), body)
            return
        gb = self._buckets(flat)
        split = len(gb.buckets) > 1 and self.dp_overlap
        if self.dp_graph_collectives:
            # one CUDA graph per step, the NCCL all-reduces captured inside it on the communication stream
            def body():
                if split:
                    grads_fn(1)
                    gb.reduce(0)
                    grads_fn(2)
                    gb.reduce(1)
                else:
                    grads_fn(0)
                    gb.reduce_all()
                gb.wait()
                eng.optim_step(net, lr)
            self._replay(eng, (tag + "_dp", lr, split), body)
            return
        if not split:
            self._replay(eng, (tag + "_grads",), lambda: grads_fn(0))
            gb.reduce_all()
        else:
            self._replay(eng, (tag + "_grads1",), lambda: grads_fn(1))
            gb.reduce(0)
            self._replay(eng, (tag + "_grads2",), lambda: grads_fn(2))
            gb.reduce(1)
        gb.wait()
        self._replay(eng, (tag + "_optim", lr), lambda: eng.optim_step(net, lr))

    def _train_disc_staged(self, eng: Engine, z, alpha=None, slot=None, snapshot=True):
        """train_disc (:376-423) on the batch already staged in the engine. slot: index of pre-staged noise
        (eng.z_all / eng.alpha_all, filled by _train_staged) instead of z / alpha."""
        self.disc.train()
        self._flat_disc.reattach_grads()
        for w in self.disc.parameters():  # observable side effect of the reference (:384-389)
            w.requires_grad = True
        for w in self.gen.parameters():
            w.requires_grad = False
        if slot is None:
            eng.z_in.copy_(z, non_blocking=True)
            eng.alpha_in.copy_(self._alpha(eng.B) if alpha is None else alpha.reshape(eng.B, 1), non_blocking=True)
            zt, at, tag = eng.z_in, eng.alpha_in, "d"
        else:
            zt, at, tag = eng.z_all[slot], eng.alpha_all[slot], f"d{slot}"
        # the reference's train_disc runs the generator in the mode it was left in (:390): eval after a generate_samples
        # call (:603) until the next train_gen (:427). (Its own captured graph: the mode is baked into the kernels.)
        gen_eval = not self.gen.training
        if gen_eval:
            tag += "e"
        self._step(eng, tag, A.NET_DISC, self._flat_disc, self._lr(self.optimizer_disc),
                   lambda ph: eng.disc_grads(zt, at, training=True, phase=ph, gen_eval=gen_eval))
        if snapshot:
            self._snapshot(eng, "d")

    def _train_gen_staged(self, eng: Engine, z, slot=None):
        """train_gen (:425-461) on the batch already staged in the engine."""
        self.gen.train()
        self._flat_gen.reattach_grads()
        for w in self.disc.parameters():  # observable side effect of the reference (:433-434)
            w.requires_grad = False
        for w in self.gen.parameters():
            w.requires_grad = True
        if slot is None:
            eng.z_in.copy_(z, non_blocking=True)
            zt, tag = eng.z_in, "g"
        else:
            zt, tag = eng.z_all[slot], f"g{slot}"
        self._step(eng, tag, A.NET_GEN, self._flat_gen, self._lr(self.optimizer_gen),
                   lambda ph: eng.gen_grads(zt, training=True, phase=ph))
        self._snapshot(eng, "g")

    def _train_staged(self, eng: Engine, zs=None, alphas=None):
        """train() (:463-477): n_critic critic steps + one generator step. The noise of all steps is drawn up
        front, in the reference's order (z, alpha, z, alpha, ..., z), straight into per-step buffers the captured
        step graphs read — no RNG / copy kernels sit between the graph replays."""
        n, B = self.n_critic, eng.B
        if not self.noise_upfront:
            for i in range(n):
                z = zs[i] if zs is not None else self._normal(B)
                self._train_disc_staged(eng, z, None if alphas is None else alphas[i])
            self._train_gen_staged(eng, zs[n] if zs is not None else self._normal(B))
            return
        eng.ensure_noise(n)
        for i in range(n + 1):
            if zs is not None:
                eng.z_all[i].copy_(zs[i], non_blocking=True)
            elif _dist() is not None and self.dp_global_noise:
                eng.z_all[i].copy_(self._normal(B))
            else:
                torch.normal(0, 1, size=(B, self.latent_dims), out=eng.z_all[i])
            if i < n:
                if alphas is not None:
                    eng.alpha_all[i].copy_(alphas[i].reshape(B, 1), non_blocking=True)
                elif _dist() is not None and self.dp_global_noise:
                    eng.alpha_all[i].copy_(self._alpha(B))
                else:
                    torch.rand(B, 1, out=eng.alpha_all[i])
        if not (self.whole_call_graph and self.use_cuda_graphs) or self.unique_graphs:
            for i in range(n):
                self._train_disc_staged(eng, None, slot=i, snapshot=(i == n - 1))
            self._train_gen_staged(eng, None, slot=n)
            return
        # the six steps (kernels, collectives, optimizer updates, the two loss read-backs) as ONE graph per call
        if self._pinned is None:
            self._pinned = {k: torch.zeros(A.STATS_COUNT, dtype=torch.float32).pin_memory() for k in ("d", "g")}

        def whole():
            self._in_capture = True
            try:
                for i in range(n):
                    self._train_disc_staged(eng, None, slot=i, snapshot=(i == n - 1))
                self._train_gen_staged(eng, None, slot=n)
            finally:
                self._in_capture = False

        self._replay(eng, ("dg_call", float(self._lr(self.optimizer_disc)), float(self._lr(self.optimizer_gen)), n,
                           self.gen.training), whole)
        # host-visible side effects of the steps that a replay does not re-execute (:384-389, :433-438)
        self.disc.train()
        self.gen.train()
        for w in self.disc.parameters():
            w.requires_grad = False
        for w in self.gen.parameters():
            w.requires_grad = True
        self._flat_disc.reattach_grads()
        self._flat_gen.reattach_grads()
        eng.stepped(A.NET_DISC)   # (a replay does not re-run the Python of the steps)
        eng.stepped(A.NET_GEN)
        self._mark("d")
        self._mark("g")

    # ---- generation over a loader (film-style batch tuples) ------------------------------------------
    GEN_CHUNK = 64   # rows per generator call of the reference's class-balanced branch (conditional_gan_concat.py:488)

    def _generate_all_film_layout(self, data_loader, num_repeats=1, balanced=False, balanced_max_oversample=5,
                                  with_site=True):
        """generate_samples_all of the scripts whose loader yields (text_embedding, gene_expression, patches,
        padding_mask, disease_type, primary_site) (multi_patch_gan_dataloader.py:48). with_site=True: the 6-tuple of
        conditional_gan_film.py:447-563 / conditional_gan_img_transformer.py; with_site=False: the 4-tuple
        (real, generated, disease types real, disease types generated) of conditional_gan_concat.py:453-552 and
        conditional_gan_attention.py:407-506.

        balanced=True is the class-balanced branch of the two 4-tuple scripts (conditional_gan_concat.py:455-526):
        one pass over the loader for the real arrays, then per disease type its rows plus up to
        balanced_max_oversample x as many re-drawn ones (np.random.choice) generated GEN_CHUNK at a time from
        `data_loader.dataset`, the result shuffled. The reference reads `dataset[idx]` once PER FIELD (five times per
        sample; each read re-draws the patch subset of an over-long case), and so does this, so that a seeded numpy
        stream gives the same rows, labels and order. Chunks shorter than GEN_CHUNK are zero-padded to it for the engine
        (one engine instead of one per class remainder); z is drawn for the real rows only, as the reference does.
        In the 6-tuple scripts that branch ends in a NameError (`all_primary_site_real`, conditional_gan_film.py:563)."""
        n_fields = 6 if with_site else 5
        if balanced and with_site:
            raise NotImplementedError("balanced=True is broken in the reference for this script "
                                      "(undefined all_primary_site_real at the return, conditional_gan_film.py:563)")
        host = lambda t: t.detach().cpu().numpy()  # noqa: E731
        if not balanced:
            real, gen, labels_real, labels_gen = [], [], [[], []], [[], []]
            for i in range(num_repeats):
                for batch in data_loader:
                    text, genes, patches, ppad = batch[:4]
                    x_real, x_gen = self.generate_samples(genes.to(self.device), text, patches, ppad)
                    gen.append(host(x_gen))
                    for k in range(n_fields - 4):
                        labels_gen[k].append(host(batch[4 + k]))
                        if i == 0:
                            labels_real[k].append(host(batch[4 + k]))
                    if i == 0:
                        real.append(host(x_real))
            out = [np.vstack(real), np.vstack(gen)]
            for k in range(n_fields - 4):
                out += [np.concatenate(labels_real[k]), np.concatenate(labels_gen[k])]
            return tuple(out)
        real, disease_real = [], []
        for batch in data_loader:
            real.append(host(batch[1].clone().to(torch.float32)))
            disease_real.append(host(batch[4]))
        real, disease_real = np.vstack(real), np.concatenate(disease_real)
        classes = np.unique(disease_real)
        counts = np.bincount(disease_real)
        most = counts.max()
        rows_of = {c: np.where(disease_real == c)[0] for c in classes}
        dataset = data_loader.dataset
        gen, disease_gen = [], []
        for _ in range(num_repeats):
            for c in classes:
                rows = rows_of[c]
                if counts[c] < most:
                    extra = min(most - counts[c], balanced_max_oversample * counts[c])
                    rows = np.concatenate((rows, np.random.choice(rows, extra, replace=extra > len(rows))))
                for s in range(0, len(rows), self.GEN_CHUNK):
                    chunk = rows[s:s + self.GEN_CHUNK]
                    # field by field per sample, in the reference's order (one dataset read per field)
                    fields = [[] for _ in range(n_fields)]
                    for r in chunk:
                        for k in range(n_fields):
                            fields[k].append(dataset[int(r)][k])
                    text, genes, patches, ppad = (torch.stack(fields[k]) for k in range(4))
                    gen.append(host(self._generate_padded(genes, text, patches, ppad, self.GEN_CHUNK)))
                    disease_gen.append(np.asarray([int(v) for v in fields[4]], dtype=np.int64))
        gen, disease_gen = np.vstack(gen), np.concatenate(disease_gen)
        order = np.arange(gen.shape[0])
        np.random.shuffle(order)
        return real, gen[order], disease_real, disease_gen[order]

    def _generate_padded(self, genes, text, patches, ppad, rows):
        """generate_samples on n <= rows samples through the engine of batch size `rows`: z ~ N(0, 1) for the n real rows
        (the reference's draw), the conditioning zero-padded; returns the n generated rows."""
        n = genes.shape[0]
        if n == rows:
            return self.generate_samples(genes.to(self.device), text, patches, ppad)[1]
        with torch.no_grad():
            self.gen.eval()
            z = torch.zeros(rows, self.latent_dims, device=self.device)
            z[:n] = torch.normal(0, 1, size=(n, self.latent_dims), device=self.device)

            def pad(t):
                t = t.to(self.device)
                out = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                out[:n] = t
                return out
            return self.gen(z, pad(text), pad(patches), pad(ppad))[:n]

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    def print_best_epoch(self, d, name="correlation"):
        idx = np.argmax(list(d.values()))
        print("Best epoch " + name + ":", list(d.keys())[idx], "score:", list(d.values())[idx])

    def evaluate_generated(self, data_real, data_gen, test_real, test_gen, nn=10, privacy=True) -> dict:
        """The GPU-computable part of the evaluation the reference's fit() runs on the arrays that
        generate_samples_all returns (src/conditional_gan_cross_attention_with_film.py:732-734, :808-811, :983-984):
        gamma_coef(test_real, test_gen), compute_prdc on the train and test pairs (the `precision` .. `coverage` and
        `*_test` entries of compute_evaluation_metrics, src/unsupervised_metrics.py:52-62) and the DCR / NNDR privacy
        scores. The sklearn classifiers, PCA and UMAP plots of the reference stay on the host and are out of scope."""
        from . import evalmetrics as em

        out = {"gamma": em.gamma_coef(test_real, test_gen)}
        train, test = em.compute_prdc(data_real, data_gen, nearest_k=nn), em.compute_prdc(test_real, test_gen, nearest_k=nn)
        for key in train:
            out[key], out[key + "_test"] = train[key], test[key]
        if privacy:
            out["dcr"] = em.dcr(data_real, data_gen, test_real)
            out["nndr"] = em.nndr(data_real, data_gen, test_real)
        return out

    def save_generated_run(self, train_data, test_data, run: int, epoch: int, evaluate: bool = True) -> dict:
        """One run of the reference's final test block (…with_film.py:786-811): generate_samples_all over the training
        and the test loader, the twelve arrays written under <results_dire>/test_<run>_epoch_<epoch+1>/ with the
        reference's file names (what UtilityEvaluatorPrimary and the privacy block read back, :962-990), and the GPU
        metrics of evaluate_generated. Returns {"folder": ..., **metrics}."""
        train_out = self.generate_samples_all(train_data)
        test_out = self.generate_samples_all(test_data)
        folder = os.path.join(self.result_dire, f"test_{run}_epoch_{epoch + 1}")
        save_generated_arrays(folder, train_out, test_out)
        out = {"folder": folder}
        if evaluate:
            out.update(self.evaluate_generated(train_out[0], train_out[1], test_out[0], test_out[1]))
        return out

    def _fit_evaluation(self, epoch, epochs, train_data, val_data, test_data, val=True, n_runs=2) -> None:
        """Evaluation block of the reference's fit() (…with_film.py:702-811), GPU-computable part. Nothing happens
        unless the caller hands fit() a validation / test loader and a results directory:
          * every freq_compute_test epochs, with val_data: generate on the training and validation loaders, record
            precision_test / recall_test (:732-734) and the gamma coefficient per epoch in `precision_scores`,
            `recall_scores`, `corr_scores` (the dicts print_best_epoch reads);
          * at the last epoch, with test_data: n_runs x save_generated_run (:786-811) -> `self.test_runs`."""
        if not (val and self.result_dire and self._is_main_rank()):
            return
        if val_data is not None and (epoch + 1) % self.freq_compute_test == 0:
            tr_out, va_out = self.generate_samples_all(train_data), self.generate_samples_all(val_data)
            m = self.evaluate_generated(tr_out[0], tr_out[1], va_out[0], va_out[1], privacy=False)
            self.precision_scores[epoch + 1] = m["precision_test"]
            self.recall_scores[epoch + 1] = m["recall_test"]
            self.corr_scores[epoch + 1] = m["gamma"]
        if test_data is not None and epoch + 1 == epochs:
            self.test_runs = [self.save_generated_run(train_data, test_data, run, epoch) for run in range(n_runs)]

    @staticmethod
    def _is_main_rank() -> bool:
        """Data-parallel replicas hold identical weights: only rank 0 writes files (checkpoints, .npy dumps)."""
        d = _dist()
        return d is None or d.get_rank() == 0

    def _save_checkpoints(self, tag: str) -> None:
        """generator_<tag>.pt / discriminator_<tag>.pt under results_dire — the reference's file names
        (…with_film.py:710-711, :743-744). Rank 0 only in data-parallel runs."""
        if not (self.result_dire and self._is_main_rank()):
            return
        torch.save(self.gen.state_dict(), os.path.join(self.result_dire, f"generator_{tag}.pt"))
        torch.save(self.disc.state_dict(), os.path.join(self.result_dire, f"discriminator_{tag}.pt"))

    def _epoch_lr_decay(self, epoch, every):
        if epoch > 0 and epoch % every == 0:
            for opt in (self.optimizer_disc, self.optimizer_gen):
                for g in opt.param_groups:
                    g["lr"] *= 0.5


# ---- the .npy layout fit() leaves behind for the utility / privacy evaluation (…with_film.py:793-806, :970-976)
GENERATED_FILES = ("data_real", "data_gen", "train_labels_real", "train_labels_gen", "train_primary_site_real",
                   "train_primary_site_gen")
GENERATED_TEST_FILES = ("test_real", "test_gen", "test_labels_real", "test_labels_gen", "test_primary_site_real",
                        "test_primary_site_gen")


def save_generated_arrays(folder: str, train_out, test_out) -> None:
    """Writes what generate_samples_all returned for the training loader (real, gen, disease type real / gen, primary
    site real / gen) and for the test loader under the reference's twelve file names (:793-806); the scripts whose
    generate_samples_all returns four arrays (no primary sites) leave the first eight
    (conditional_gan_concat.py:760-767, conditional_gan_attention.py:639-646)."""
    os.makedirs(folder, exist_ok=True)
    for names, arrays in ((GENERATED_FILES, train_out), (GENERATED_TEST_FILES, test_out)):
        if len(arrays) == 4:
            names = names[:4]
        if len(arrays) != len(names):
            raise ValueError(f"expected the 6- or 4-tuple of generate_samples_all, got {len(arrays)} arrays")
        for name, a in zip(names, arrays):
            with open(os.path.join(folder, name + ".npy"), "wb") as f:
                np.save(f, np.asarray(a))


def load_generated_arrays(folder: str) -> dict:
    """load_data of the reference's privacy block (:970-976)."""
    return {k: np.load(os.path.join(folder, k + ".npy")) for k in ("data_real", "data_gen", "test_real", "test_gen")}


def privacy_report(output_path: str) -> dict:
    """DCR / NNDR over every <output_path>/test_* folder, mean and std (:978-995), on the GPU kernels."""
    from glob import glob

    from . import evalmetrics as em

    dcr_scores, nndr_scores = [], []
    for folder in sorted(glob(os.path.join(output_path, "test_*"))):
        data = load_generated_arrays(folder)
        dcr_scores.append(em.dcr(data["data_real"], data["data_gen"], data["test_real"]))
        nndr_scores.append(em.nndr(data["data_real"], data["data_gen"], data["test_real"]))
    return {"dcr": dcr_scores, "nndr": nndr_scores, "mean_dcr": float(np.mean(dcr_scores)),
            "std_dcr": float(np.std(dcr_scores)), "mean_nndr": float(np.mean(nndr_scores)),
            "std_nndr": float(np.std(nndr_scores))}
