"""Data-parallel gradient synchronisation for the flat fp32 gradient buffer of FusedClipAdam.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The image batch is the only sharded dimension; the
single collective of a step is a bucketed all-reduce (average) of the student's gradients.  FusedClipAdam lays parameters
out in forward-execution order, so backward finalises the flat gradient buffer back-to-front: `GradBucketer.notify(offset)`
(called as backward passes stage boundaries) launches the all-reduce of every bucket that lies completely behind the
frontier on a side stream, overlapping communication with the remaining backward kernels.  The EMA teacher update stays
rank-local: every rank applies the same averaged gradients, hence holds bit-identical student (and teacher) weights.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradBucketer:
    def __init__(self, flat_grad: torch.Tensor, bucket_bytes: int = 32 << 20, group=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat_grad.numel()
        per = max(1, bucket_bytes // 4)
        # buckets are enumerated from the END of the buffer (first to become final during backward)
        self.buckets: List[tuple] = []
        end = n
        while end > 0:
            start = max(0, end - per)
            self.buckets.append((start, end))
            end = start
        self._next = 0
        self._works = []
        self._nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._side: Optional[torch.cuda.Stream] = None
        if flat_grad.is_cuda:
            self._side = torch.cuda.Stream(device=flat_grad.device)

    def begin_step(self) -> None:
        self._next = 0
        self._works = []

    def _launch(self, bucket) -> None:
        a, b = bucket
        view = self.flat[a:b]
        # NCCL averages in the collective itself; gloo (CPU tests) has no AVG, so sums are scaled in finish()
        op = dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM
        if self._side is not None:
            from . import ops
            ops.join_wgrad()  # weight gradients of this bucket may still be in flight on the wgrad stream
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                w = dist.all_reduce(view, op=op, group=self.group, async_op=True)
        else:
            w = dist.all_reduce(view, op=op, group=self.group, async_op=True)
        self._works.append(w)

    def notify(self, frontier_offset: int) -> None:
        """All gradients at flat offsets >= frontier_offset are final: reduce the buckets that start at or after it."""
        if self.world == 1:
            return
        while self._next < len(self.buckets) and self.buckets[self._next][0] >= frontier_offset:
            self._launch(self.buckets[self._next])
            self._next += 1

    def finish(self) -> None:
        """Reduce whatever is left, wait, and turn the sums into means."""
        if self.world == 1:
            return
        self.notify(0)
        for w in self._works:
            w.wait()
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        if not self._nccl:
            self.flat.mul_(1.0 / self.world)
        self._works = []


def broadcast_flat(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """Make every rank start from rank-0's parameters (called once after construction)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)


def attach(model_wrapper, bucket_bytes: int = 32 << 20, group=None) -> GradBucketer:
    """Wire a SegFormerModel wrapper for data-parallel training: broadcast rank-0 parameters, build the bucketer over the
    optimiser's flat gradient buffer and tell the encoder at which flat offsets each stage's gradients become final."""
    opt = model_wrapper.optimizer
    broadcast_flat(opt.flat_p, 0, group)
    for buf in model_wrapper.model.buffers():
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(buf, src=0, group=group)
    bucketer = GradBucketer(opt.flat_g, bucket_bytes, group)
    offset_of = {id(p): o for p, o in zip(opt._params, opt._offsets)}
    enc = model_wrapper.model.segformer.encoder
    stage_offsets = []
    for blk in enc.block:
        offs = [offset_of[id(p)] for p in blk.parameters() if id(p) in offset_of]
        stage_offsets.append(min(offs) if offs else 0)
    enc._ddp_bucketer = bucketer
    enc._ddp_stage_offsets = stage_offsets
    return bucketer
