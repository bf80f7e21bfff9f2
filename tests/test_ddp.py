"""Host-side data-parallel logic on CPU: gloo, world_size 2.

Covers gemmgan_b200/ddp.py: the bucket plan over the flat gradient layout, the asynchronous bucketed
all-reduce(mean) and the global-noise slicing that makes N ranks x B identical to 1 rank x N*B.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gemmgan_b200 import _abi_decl as A
from gemmgan_b200.ddp import Bucket, GradBuckets, global_noise, plan_buckets

TRUNK = (A.P_TR0_W, A.P_TR0_B, A.P_TR1_W, A.P_TR1_B, A.P_FIN_W, A.P_FIN_B)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_buckets_trunk_first():
    offsets = {A.P_FILM_W: 0, A.P_TEXT_W: 640, A.P_TR0_W: 1280, A.P_TR0_B: 6400, A.P_FIN_W: 6464, A.P_FIN_B: 9000}
    b = plan_buckets(offsets, 9064, TRUNK)
    assert [(x.name, x.start, x.stop) for x in b] == [("trunk", 1280, 9064), ("tower", 0, 1280)]
    # vanilla: trunk tensors only -> one bucket covering everything
    b = plan_buckets({A.P_TR0_W: 0, A.P_FIN_B: 512}, 576, TRUNK)
    assert [(x.start, x.stop) for x in b] == [(0, 576)]


def test_plan_buckets_rejects_tower_behind_trunk():
    with pytest.raises(AssertionError):
        plan_buckets({A.P_TR0_W: 0, A.P_FILM_W: 512}, 1024, TRUNK)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
        gb = GradBuckets(flat, [Bucket("trunk", 600, n), Bucket("tower", 0, 600)])
        assert gb.bytes() == [1600, 2400]
        gb.reduce(0)              # trunk bucket starts while "the tower backward" still writes the other slice
        flat[:600] += 1.0
        gb.reduce(1)
        gb.wait()
        mean_scale = sum(r + 1 for r in range(world)) / world
        want = torch.arange(n, dtype=torch.float32) * mean_scale
        want[:600] += 1.0
        assert torch.allclose(flat, want)
        # global noise: every rank draws the global batch from the shared seed and keeps its rows
        g = torch.Generator().manual_seed(42)
        mine = global_noise(lambda k: torch.randn(k, 8, generator=g), per_rank=4)
        full = torch.randn(world * 4, 8, generator=torch.Generator().manual_seed(42))
        assert torch.equal(mine, full[rank * 4:(rank + 1) * 4])
        # mean of per-rank batch-mean gradients == gradient of the global batch mean (equal per-rank batches)
        w = torch.ones(8, requires_grad=True)
        (mine @ w).pow(2).mean().backward()
        gr = w.grad.clone()
        dist.all_reduce(gr)
        gr /= world
        w2 = torch.ones(8, requires_grad=True)
        (full @ w2).pow(2).mean().backward()
        assert torch.allclose(gr, w2.grad, atol=1e-6)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
