"""The data-parallel TRAINER on the CPU suite: two gloo processes, each a drop-in `WGAN_GP`
(gemmgan_b200/trainer.py) on the host-emulated engine (tests/host_trainer.py), against one process on the global batch.

Covers what tests/test_dp_emulated.py (engine + buckets driven by hand) does not: the trainer's own data-parallel code —
rank 0's initial weights broadcast to every replica whatever each rank's torch seed was, the staged gradient buckets
issued from `_step_inner`, z / alpha drawn for the global batch and sliced (`dp_global_noise`), the optimizer steps of
two train() calls with the replicas staying bit-identical, rank 0 evaluating alone in between (a new engine without a
collective), and the refusal of unequal per-rank batches."""
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


def _load(lib_path):
    import host_trainer
    from gemmgan_b200 import _abi_decl as A

    L = C.CDLL(lib_path)
    L.gg_last_error.restype = C.c_char_p
    A.declare(L)
    host_trainer.apply(setattr, L)


def _trainer(variant, seed):
    c = CFG
    H, G = c["hidden"], c["G"]
    torch.manual_seed(seed)
    if variant == "vanilla":
        import vanilla_gan_unconditional as m
        t = m.WGAN_GP_nocond(input_dims=G, latent_dims=c["latent"], vocab_sizes=[], generator_dims=[H, H, G],
                             discriminator_dims=[H, H, 1], optimizer="adam")
        t.build_WGAN_GP_nocond()
    else:
        import conditional_gan_cross_attention_with_film as m
        t = m.WGAN_GP(input_dims=G, latent_dims=c["latent"], embedding_dims=c["embed"], generator_dims=[H, H, G],
                      discriminator_dims=[H, H, 1], text_embedding_dims=c["text_dim"],
                      patches_embedding_dims=c["patch_dim"], optimizer="adam")
        t.build_WGAN_GP()
    t.dp_global_noise = True
    t.init_train()
    return t


def _batch(variant, n):
    from oracle import restated    # synthetic batch helper only

    c = CFG
    return restated.synthetic_batch(variant, n, c["G"], c["P"], c["T"], seed=5, ragged=True, text_dim=c["text_dim"],
                                    patch_dim=c["patch_dim"])


def _train(t, variant, x, cond, rows=slice(None)):
    if variant == "vanilla":
        t.train(x[rows])
    else:
        patches, ppad, text, tpad = cond
        t.train(x[rows], text[rows], tpad[rows], patches[rows], ppad[rows])


def _state(t):
    return {"gen": {k: v.clone() for k, v in t.gen.state_dict().items()},
            "disc": {k: v.clone() for k, v in t.disc.state_dict().items()},
            "d": t.d_batch_loss.copy(), "g": t.g_batch_loss.copy()}


def _worker(rank, world, port, lib_path, out_dir, variant):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _load(lib_path)
        B = CFG["B"]
        t = _trainer(variant, seed=100 + rank)          # every rank its own seed: rank 0's weights must win
        init = {k: v.clone() for k, v in t.disc.state_dict().items()}
        x, cond = _batch(variant, world * B)
        torch.manual_seed(77)                           # shared noise seed (global z / alpha drawn on every rank)
        _train(t, variant, x, cond, slice(rank * B, (rank + 1) * B))
        if rank == 0:
            # rank 0 evaluates alone between epochs (fit()'s evaluation block): a new batch size means a new engine,
            # created without any collective, or the other ranks' next gradient all-reduce would pair up with it
            xv, cv = _batch(variant, 3)
            if variant == "vanilla":
                t.generate_samples(xv)
            else:
                t.generate_samples(xv, cv[2], cv[3], cv[0], cv[1])
        torch.manual_seed(78)
        _train(t, variant, x, cond, slice(rank * B, (rank + 1) * B))
        out = _state(t)
        out["init_disc"] = init
        torch.save(out, os.path.join(out_dir, f"rank{rank}.pt"))
        # a last partial batch that differs between the ranks is refused, not averaged with equal weights
        try:
            _train(t, variant, x, cond, slice(0, B - 1 - rank))
            refused = False
        except ValueError as e:
            refused = "equal per-rank batch sizes" in str(e)
        torch.save({"refused": refused}, os.path.join(out_dir, f"refused{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _cosine(before, a, b):
    ua = torch.cat([(v - before[k]).flatten() for k, v in a.items()])
    ub = torch.cat([(v - before[k]).flatten() for k, v in b.items()])
    return float(ua @ ub / (ua.norm() * ub.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("variant", ["paper", "vanilla"])
def test_two_trainer_ranks_equal_one_trainer_on_the_global_batch(variant, tmp_path_factory, tmp_path):
    out = tmp_path_factory.mktemp("cuda_emu")
    emu_build.build("engine", out, cudart=True)
    lib_path = os.path.join(str(out), "libengine_emu.so")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), lib_path, str(tmp_path), variant), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world))
    # replicas: same start (rank 0's weights) and bit-identical after six optimizer steps
    for net in ("gen", "disc"):
        for k in r0[net]:
            assert torch.equal(r0[net][k], r1[net][k]), (net, k)
    for k in r0["init_disc"]:
        assert torch.equal(r0["init_disc"][k], r1["init_disc"][k]), k
    assert all(torch.load(tmp_path / f"refused{r}.pt", weights_only=False)["refused"] for r in range(world))

    # one process on the whole batch, same initial weights (seed of rank 0), same global noise
    _load(lib_path)
    t = _trainer(variant, seed=100)
    for k, v in t.disc.state_dict().items():
        assert torch.equal(v, r0["init_disc"][k]), k
    before = {"gen": {k: v.clone() for k, v in t.gen.state_dict().items()},
              "disc": {k: v.clone() for k, v in t.disc.state_dict().items()}}
    x, cond = _batch(variant, world * CFG["B"])
    torch.manual_seed(77)
    _train(t, variant, x, cond)
    torch.manual_seed(78)
    _train(t, variant, x, cond)
    one = _state(t)
    # losses: mean over ranks of the per-rank batch means = the global batch mean
    d_mean = 0.5 * (r0["d"] + r1["d"])
    assert abs(d_mean - one["d"]).max() <= 2e-2 * max(1.0, abs(one["d"]).max()), (d_mean, one["d"])
    assert abs(0.5 * (r0["g"] + r1["g"]) - one["g"]).max() <= 2e-2 * max(1.0, abs(one["g"]).max())
    # weights: the same update direction (Adam's first steps are lr * sign(g): compare as update vectors)
    cd = _cosine(before["disc"], r0["disc"], one["disc"])
    cg = _cosine(before["gen"], r0["gen"], one["gen"])
    print(f"{variant}: 2 ranks vs 1 rank update cosine: critic {cd:.4f}, generator {cg:.4f}")
    assert cd > 0.9 and cg > 0.9, (cd, cg)
