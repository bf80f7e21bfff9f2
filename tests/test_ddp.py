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
from gemmgan_b200.ddp import Bucket, GradBuckets, global_noise, plan_buckets, plan_stage_buckets

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


def test_plan_stage_buckets_paper_and_film_layouts():
    cross = tuple(range(A.P_P2T_IN_W, A.P_T2P_OUT_B + 1))
    # paper-like layout: every slot present, 64 elements each
    offsets = {s: 64 * i for i, s in enumerate(range(A.NSLOTS))}
    plan = plan_stage_buckets(offsets, 64 * A.NSLOTS, TRUNK, A.P_LAYER0, A.L_COUNT, 2, cross)
    assert [st for st, _ in plan] == [-1, 0, 1, 2, 3]
    assert [b.name for _, b in plan] == ["trunk", "cross", "layer1", "layer0", "embed"]
    spans = sorted((b.start, b.stop) for _, b in plan)
    assert spans[0][0] == 0 and spans[-1][1] == 64 * A.NSLOTS
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, disjoint, complete
    by = {b.name: b for _, b in plan}
    assert by["layer1"].start == offsets[A.P_LAYER0 + A.L_COUNT] and by["cross"].start == offsets[A.P_P2T_IN_W]
    # film-like layout: no text encoder, no cross-attention, bias-free layers (only the weight slots present)
    slots = [A.P_FILM_W, A.P_FILM_B, A.P_PATCH_W, A.P_PATCH_B, A.P_CLS]
    for layer in range(2):
        base = A.P_LAYER0 + A.L_COUNT * layer
        slots += [base + A.L_IN_W, base + A.L_OUT_W, base + A.L_FF1_W, base + A.L_FF2_W, base + A.L_N1_W, base + A.L_N2_W]
    slots += list(TRUNK)
    offsets = {s: 128 * i for i, s in enumerate(sorted(slots))}
    plan = plan_stage_buckets(offsets, 128 * len(slots), TRUNK, A.P_LAYER0, A.L_COUNT, 2, cross)
    assert [st for st, _ in plan] == [-1, 1, 2, 3]
    assert sum(b.stop - b.start for _, b in plan) == 128 * len(slots)


def test_plan_stage_buckets_label_and_concat_layouts():
    """Trunk-plus-one-tensor-group variants: the label-conditioned baseline (two embedding tables in the EMB0 / EMB1
    slots) and the concat model (one encoder): trunk bucket first, everything else in the stage-3 'embed' bucket,
    which the staged step reduces after the stage that produced it (stage 0)."""
    cross = tuple(range(A.P_P2T_IN_W, A.P_T2P_OUT_B + 1))
    for slots in ([A.P_EMB0, A.P_EMB1], [A.P_TEXT_W, A.P_TEXT_B]):
        offsets, off = {}, 0
        for s_ in sorted(slots) + list(TRUNK):
            offsets[s_] = off
            off += 192
        plan = plan_stage_buckets(offsets, off, TRUNK, A.P_LAYER0, A.L_COUNT, 2, cross)
        assert [(st, b.name, b.start, b.stop) for st, b in plan] == [(-1, "trunk", 384, off), (3, "embed", 0, 384)]


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
