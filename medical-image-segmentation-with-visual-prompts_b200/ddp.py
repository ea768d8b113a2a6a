"""Batch-parallel gradient exchange for the prompted Swin blocks ("DDP-lite").

The path shards by batch: one process per GPU, full replica, no data-path collective; the only
exchange step is ONE all-reduce (sum -> / world) of the gradients per optimiser step over NCCL/NVLink.  The
reference's trainers call parameter-group accessors on the raw module and run 2-3 forwards per backward
(multi_view.py:59,135-136), so instead of wrapping the module this hooks the parameters:
`register_post_accumulate_grad_hook` marks gradients as ready, full buckets are all-reduced asynchronously while
the rest of backward still runs, and `finish()` waits and writes the averaged gradients back.
Works with any torch.distributed backend (nccl on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class BucketedGradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None, overlap: bool = True):
        """overlap=True: a bucket's all-reduce starts (in bucket order) as soon as its last gradient is accumulated --
        exactly ONE backward per finish().  overlap=False: every collective is issued from finish(), any number of
        backward passes (gradient accumulation) in between."""
        self.group = group
        self.overlap = overlap
        self._next = 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        # reverse order ~ the order in which backward produces gradients
        self.buckets: List[List[int]] = []
        cur, size = [], 0
        for i in reversed(range(len(self.params))):
            n = self.params[i].numel() * 4
            if cur and size + n > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(i)
            size += n
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {i: b for b, idxs in enumerate(self.buckets) for i in idxs}
        self._ready = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._handles = []
        if self.world > 1:
            for i, p in enumerate(self.params):
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, i):
        def hook(_param):
            b = self.bucket_of[i]
            self._ready[b] += 1
            if self._ready[b] > len(self.buckets[b]):
                # a second backward before finish() (gradient accumulation): the bucket's all-reduce of the first
                # backward is already in flight and would overwrite the accumulated gradients in finish()
                raise RuntimeError("BucketedGradSync: backward() ran twice before finish(); with gradient accumulation "
                                   "construct it with overlap=False (collectives are then issued from finish() only)")
            # Buckets are launched strictly in bucket order on every rank: a bucket whose hooks all fired waits for its
            # predecessors (ranks with different sets of unused parameters would otherwise issue the collectives in
            # different orders, which mismatches or hangs NCCL)
            if self.overlap:
                while self._next < len(self.buckets) and self._ready[self._next] == len(self.buckets[self._next]):
                    self._launch(self._next)
                    self._next += 1
        return hook

    def _launch(self, b):
        grads = []
        for i in self.buckets[b]:
            p = self.params[i]
            grads.append((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float())
        flat = torch.cat(grads)
        self._flat[b] = flat
        self._work[b] = dist.all_reduce(flat, group=self.group, async_op=True)

    def finish(self):
        """Call after backward (once per optimiser step): launches buckets whose hooks did not all fire
        (parameters without gradient this step), waits, averages and scatters back into .grad."""
        if self.world == 1:
            return
        for b in range(self._next, len(self.buckets)):     # in bucket order, like the hooks
            self._launch(b)
        self._next = 0
        for b, idxs in enumerate(self.buckets):
            self._work[b].wait()
            flat = self._flat[b].div_(self.world)
            o = 0
            for i in idxs:
                p = self.params[i]
                n = p.numel()
                g = flat[o:o + n].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                o += n
            self._ready[b], self._work[b], self._flat[b] = 0, None, None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def allreduce_gradients_flat(params: Iterable[torch.nn.Parameter], group=None):
    """Single flat-bucket variant (what bench.py uses: 15.5 MB at cfg2, latency-bound)."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world == 1:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
