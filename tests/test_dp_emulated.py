"""N ranks x B == 1 rank x N*B on the CPU suite: two gloo processes, each with the host-emulated engine
(tests/cuda_emu/emu_engine.cpp) on its half of the global batch and the staged gradient buckets of gemmgan_b200/ddp.py
(all-reduce(mean) per backward stage, as the data-parallel trainer issues them), against ONE engine on the whole batch.
z / alpha are drawn for the global batch and sliced (ddp.global_noise), the data rows are [rank*B, (rank+1)*B). On the
B200 the same comparison is tests/gpu_dp_parity.py over NCCL (profiles/r01_dp_parity_n2.log)."""
import contextlib
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_build

CFG = dict(B=4, G=203, P=5, T=3, embed=32, hidden=32, latent=16, text_dim=24, patch_dim=32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(lib_path):
    """Points gemmgan_b200.runtime at the emulated library (what the `rt` fixture of test_engine_emulated.py does)."""
    from gemmgan_b200 import _abi_decl as A
    from gemmgan_b200 import _lib, runtime

    L = C.CDLL(lib_path)
    L.gg_last_error.restype = C.c_char_p
    A.declare(L)
    _lib.lib = lambda: L
    _lib.require_device = lambda dev=0: None
    runtime._stream = lambda: None
    torch.cuda.device = lambda d: contextlib.nullcontext()
    return runtime


def _engine(rt, B):
    import conditional_gan_cross_attention_with_film as m
    from gemmgan_b200 import _lib

    c = CFG
    torch.manual_seed(11)
    H, G = c["hidden"], c["G"]
    gen, disc = m.WGAN_GP_model(c["latent"], G, c["embed"], [H, H, G], [H, H, 1], c["text_dim"], c["patch_dim"], 0.0, False)
    dev = torch.device("cpu")
    fg, fd = rt.FlatNet(gen, dev, "adam"), rt.FlatNet(disc, dev, "adam")
    eng = rt.Engine(variant="paper", B=B, G=G, L=c["latent"], gen=fg, disc=fd, slope=0.0, dropout_p=0.0, gp_weight=10.0,
                    clip_d=10.0, clip_g=2.0, optimizer="adam", gemm_impl=_lib.IMPL_SIMT_F32, device=dev, E=c["embed"],
                    H=H, Dt=c["text_dim"], Dp=c["patch_dim"], P=c["P"], T=c["T"], tower_bias=True)
    eng.set_lanes(False)
    return eng


def _batch(world):
    from oracle import restated    # synthetic batch helper only (tests may use the oracle package)

    c = CFG
    return restated.synthetic_batch("paper", world * c["B"], c["G"], c["P"], c["T"], seed=5, ragged=True,
                                    text_dim=c["text_dim"], patch_dim=c["patch_dim"])


def _noise(n):
    g = torch.Generator().manual_seed(42)
    return torch.randn(n, CFG["latent"], generator=g), torch.rand(n, 1, generator=g)


def _worker(rank, world, port, lib_path, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gemmgan_b200 import _abi_decl as A
        from gemmgan_b200 import ddp

        rt = _setup(lib_path)
        B = CFG["B"]
        eng = _engine(rt, B)
        x, (patches, ppad, text, tpad) = _batch(world)
        rows = slice(rank * B, (rank + 1) * B)
        eng.set_batch(genes=x[rows], patches=patches[rows], patch_pad=ppad[rows], text=text[rows], text_pad=tpad[rows])
        zs, alphas = _noise(world * B)
        z = ddp.global_noise(lambda n: zs[:n], B)
        alpha = ddp.global_noise(lambda n: alphas[:n], B)
        flat = eng.disc
        cross = tuple(range(A.P_P2T_IN_W, A.P_T2P_OUT_B + 1))
        trunk = (A.P_TR0_W, A.P_TR0_B, A.P_TR1_W, A.P_TR1_B, A.P_FIN_W, A.P_FIN_B)
        plan = ddp.plan_stage_buckets(flat.offsets, flat.n_used, trunk, A.P_LAYER0, A.L_COUNT, 2, cross)
        gb = ddp.GradBuckets(flat.grads, [b for _, b in plan])
        stages = {st: i for i, (st, _) in enumerate(plan)}
        nj = A.PHASE_NO_JOIN
        eng.disc_grads(z, alpha, training=True, phase=1 | nj)            # forward + trunk backward
        gb.reduce(stages[-1])
        for st in range(4):                                              # head, two encoder layers, embedding tail
            eng.disc_grads(z, alpha, training=True, phase=(A.PHASE_STAGE0 + st) | (0 if st == 3 else nj))
            if st in stages:
                gb.reduce(stages[st])
        gb.wait()
        eng.optim_step(A.NET_DISC, 5e-4)
        torch.save({"grads": flat.grads.clone(), "params": flat.params.clone(), "stats": eng.stats.clone()},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_the_global_batch(tmp_path_factory, tmp_path):
    out = tmp_path_factory.mktemp("cuda_emu")
    emu_build.build("engine", out, cudart=True)
    lib_path = os.path.join(str(out), "libengine_emu.so")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), lib_path, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert torch.equal(r0["grads"], r1["grads"]) and torch.equal(r0["params"], r1["params"])   # replicas stay identical

    # one process, the whole batch
    from gemmgan_b200 import _abi_decl as A
    rt = _setup(lib_path)
    eng = _engine(rt, world * CFG["B"])
    x, (patches, ppad, text, tpad) = _batch(world)
    eng.set_batch(genes=x, patches=patches, patch_pad=ppad, text=text, text_pad=tpad)
    z, alpha = _noise(world * CFG["B"])
    eng.disc_grads(z, alpha, training=True)
    # losses: mean over ranks of the per-rank batch means = the global batch mean
    st = 0.5 * (r0["stats"] + r1["stats"])
    for k in (A.STAT_LOSS_REAL, A.STAT_LOSS_FAKE, A.STAT_GP):
        assert st[k].item() == pytest.approx(eng.stats[k].item(), rel=2e-3, abs=2e-4), k
    eng.optim_step(A.NET_DISC, 5e-4)
    g1, gN = eng.disc.grads, r0["grads"]
    # same arithmetic, different batch split: bf16 roundings of a few intermediates differ -> ~1e-2 Frobenius
    assert ((gN - g1).norm() / g1.norm()).item() < 2e-2
