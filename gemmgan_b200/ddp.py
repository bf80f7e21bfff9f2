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

    def reduce(self, i: int) -> None:
        """Start the all-reduce of bucket i; everything enqueued so far on the current stream produces it."""
        d = dist_or_none()
        if d is None:
            return
        v = self.views[i]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            self.comm_stream.wait_event(ready)
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
