"""Data-parallel plumbing for the WGAN-GP step (absent in the reference, which is single-process `cuda:0`:
src/conditional_gan_cross_attention_with_film.py:301).

The step shards by batch: every loss is a batch mean (:34, :374) and samples are independent, so with equal
per-rank batches the global gradient is the mean of the per-rank gradients. One process per GPU holds a full
replica; the only exchange is an all-reduce(mean) of the flat fp32 gradient buffer between the backward and
the optimizer kernel, six times per `train()` (5 critic steps + 1 generator step).

* `GradBuckets` cuts the flat gradient buffer into contiguous buckets in the order the hand-written backward
  finishes them (the critic/generator trunk first — over half of each net is one trunk matrix — then the
  fusion tower) and reduces each bucket on a communication stream as soon as its producer has been enqueued,
  so that the trunk bucket's all-reduce overlaps the tower backward. NCCL over NVLink/NVSwitch on GPUs; the
  same code runs with gloo on CPU tensors (tests).
* `global_noise` draws z / alpha for the GLOBAL batch from the shared seed on every rank and returns this
  rank's rows, which makes N ranks x B identical to 1 rank x N*B (parity tests).

Nothing here touches model math; PyTorch is only the transport (`torch.distributed`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """One process per GPU: run this process (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA
    node the GPU hangs off. With 8 ranks each copying ~110 MB of fp32 batch per step, host buffers that all sit on one
    node make every other rank's H2D copies cross the socket interconnect. Reads sysfs only (no libnuma); returns the
    node, or None when the platform does not expose one (then nothing changes). Call before allocating pinned memory."""
    import os

    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def dist_or_none():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def global_noise(draw: Callable[[int], torch.Tensor], per_rank: int, rank: Optional[int] = None,
                 world: Optional[int] = None) -> torch.Tensor:
    """`draw(n)` returns n rows of noise from the shared generator; every rank draws the global batch and
    keeps rows [rank*per_rank, (rank+1)*per_rank)."""
    d = dist_or_none()
    if world is None:
        world = d.get_world_size() if d is not None else 1
    if rank is None:
        rank = d.get_rank() if d is not None else 0
    full = draw(world * per_rank)
    return full[rank * per_rank:(rank + 1) * per_rank].contiguous()


@dataclass
class Bucket:
    name: str
    start: int   # element offsets into the flat gradient buffer
    stop: int


def plan_buckets(offsets: dict, n_used: int, first_slots: Sequence[int], names=("trunk", "tower")) -> List[Bucket]:
    """Two buckets from the flat layout: `first_slots` (the trunk tensors, whose gradients the backward
    finishes first) must form one contiguous range at the end of the buffer (slots are laid out in
    increasing slot order and the trunk slots are the highest); everything before it is the second bucket.
    Returns them in completion order."""
    present = sorted(offsets[s] for s in first_slots if s in offsets)
    if not present:
        return [Bucket(names[1], 0, n_used)]
    cut = present[0]
    others = [o for s, o in offsets.items() if s not in first_slots]
    assert all(o < cut for o in others), "trunk tensors must sit behind every tower tensor in the flat buffer"
    out = [Bucket(names[0], cut, n_used)]
    if cut > 0:
        out.append(Bucket(names[1], 0, cut))
    return out


def plan_stage_buckets(offsets: dict, n_used: int, trunk_slots: Sequence[int], layer0: int, layer_count: int,
                       n_layers: int, cross_slots: Sequence[int]) -> List[Tuple[int, Bucket]]:
    """Buckets in the order the staged backward finishes them, as (stage, bucket) pairs: stage -1 = trunk (after
    phase 1), 0 = cross-attention tail, 1 .. n_layers = encoder layers from the last to the first, n_layers + 1 =
    the embedding tensors (slots below the first layer: FiLM / text / patch encoders, CLS). Slots are laid out in
    increasing slot order, so each group is one contiguous range; groups a variant does not have are skipped."""
    def span(slots):
        present = sorted(offsets[s] for s in slots if s in offsets)
        return (present[0], present[-1]) if present else None

    starts = sorted(set(offsets.values())) + [n_used]

    def end_of(last_start):
        return starts[starts.index(last_start) + 1]

    out: List[Tuple[int, Bucket]] = []
    sp = span(trunk_slots)
    if sp:
        out.append((-1, Bucket("trunk", sp[0], end_of(sp[1]))))
    sp = span(cross_slots)
    if sp:
        out.append((0, Bucket("cross", sp[0], end_of(sp[1]))))
    for k in range(1, n_layers + 1):
        layer = n_layers - k
        sp = span(range(layer0 + layer_count * layer, layer0 + layer_count * (layer + 1)))
        if sp:
            out.append((k, Bucket(f"layer{layer}", sp[0], end_of(sp[1]))))
    sp = span(range(0, layer0))
    if sp:
        out.append((n_layers + 1, Bucket("embed", sp[0], end_of(sp[1]))))
    covered = sum(b.stop - b.start for _, b in out)
    assert covered == n_used, f"stage buckets cover {covered} of {n_used} gradient elements"
    return out


class GradBuckets:
    """All-reduce(mean) of contiguous slices of one flat gradient tensor, asynchronously.

    On CUDA tensors the collective is enqueued on `comm_stream` after an event recorded on the producing
    stream, and `wait()` makes the current stream wait for it (no host synchronisation). On CPU tensors
    (gloo) the async work handles are kept and `wait()` blocks on them."""

    def __init__(self, flat: torch.Tensor, buckets: List[Bucket], group=None):
        self.flat = flat
        self.buckets = buckets
        self.group = group
        self.views = [flat[b.start:b.stop] for b in buckets]
        self.cuda = flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat.device, priority=-1) if self.cuda else None
        self._pending: List[Tuple[object, object]] = []

    def bytes(self) -> List[int]:
        return [v.numel() * v.element_size() for v in self.views]

    def reduce(self, i: int, also_wait: Optional[Callable[[object], None]] = None) -> None:
        """Start the all-reduce of bucket i; everything enqueued so far on the current stream produces it.
        also_wait(comm_stream): extra producers the communication stream has to wait for (the engine's side lanes)."""
        d = dist_or_none()
        if d is None:
            return
        v = self.views[i]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            self.comm_stream.wait_event(ready)
            if also_wait is not None:
                also_wait(self.comm_stream)
            with torch.cuda.stream(self.comm_stream):
                d.all_reduce(v, op=d.ReduceOp.AVG, group=self.group)
                done = torch.cuda.Event()
                done.record()
            self._pending.append((None, done))
        else:
            # gloo has no AVG: sum, then scale on completion
            w = d.all_reduce(v, op=d.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((w, v))

    def reduce_all(self) -> None:
        for i in range(len(self.buckets)):
            self.reduce(i)

    def wait(self) -> None:
        d = dist_or_none()
        for work, x in self._pending:
            if work is None:
                torch.cuda.current_stream().wait_event(x)
            else:
                work.wait()
                x.div_(d.get_world_size(self.group))
        self._pending.clear()
