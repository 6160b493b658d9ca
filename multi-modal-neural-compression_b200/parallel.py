"""Data parallelism over the image batch: one process per GPU, gradients averaged with NCCL all-reduces that overlap
the backward pass.

The reference has no parallelism at all (its Trainer is pinned to a single device,
/root/reference/src/train.py:288-295).  Every op on the rate path is per image, parameters are
replicated, so the path shards by batch with a single exchange step in training: the mean of the main-parameter
gradients (batch-mean semantics of mtc.py:240 / 291).  Eval, compress and decompress need no collective.

Gradients live in ONE flat fp32 buffer whose slices are the `.grad` views of the parameters, laid out in reverse
registration order (roughly the order backward produces them) and cut into a few contiguous buckets.  A
post-accumulate hook per parameter counts arrivals; the moment a bucket is complete its all-reduce (average) is
issued asynchronously, so the exchange of the output heads' gradients runs on NVLink while the input heads are still
back-propagating.  The task heads back-propagate on their own CUDA streams (compressors._run_heads): the issuing
stream first waits on every stream that contributed to the bucket.  `grad_sync()` — called by the training step
between backward() and optimizer.step() — issues whatever is left and waits for all of it.
The `quantiles` parameters are left out: their gradient comes from the data-independent auxiliary loss
(/root/reference/src/models/multi_task_compressor.py:456-462) and is identical on every rank.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Re-homes the gradients of `params` into one contiguous buffer (in reverse registration order, i.e. roughly
    the order backward produces them), cut into `n_buckets` contiguous slices of similar size."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 1):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        n_buckets = max(1, min(int(n_buckets), len(self.params)))
        target = (total + n_buckets - 1) // n_buckets
        self.bucket_of = {}          # id(param) -> bucket index
        self.bounds = []             # [start, end) element ranges of the buckets
        self.counts = []             # parameters per bucket
        off, start, cur, count = 0, 0, 0, 0
        for p in reversed(self.params):
            n = p.numel()
            if count and off - start + n > target and cur < n_buckets - 1:  # close the bucket before it overflows
                self.bounds.append((start, off))
                self.counts.append(count)
                start, cur, count = off, cur + 1, 0
            p.grad = self.flat[off:off + n].view_as(p)
            self.bucket_of[id(p)] = cur
            off += n
            count += 1
        self.bounds.append((start, off))
        self.counts.append(count)
        self.slices = [self.flat[a:b] for a, b in self.bounds]

    def zero(self) -> None:
        self.flat.zero_()  # one memset instead of one kernel per parameter

    def check_views(self) -> bool:
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


class DataParallel:
    """Wraps a compressor: its training step zeroes the flat bucket, back-propagates (bucket all-reduces start as
    the buckets fill) and calls `grad_sync`, which waits for the exchange before the optimizer step."""

    def __init__(self, compressor, process_group: Optional[dist.ProcessGroup] = None, n_buckets: int = 6,
                 overlap: bool = True):
        self.module = compressor
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.n_buckets, self.overlap = int(n_buckets), bool(overlap)
        self._hooks = []
        self.rebuild_bucket()
        compressor.grad_sync = self.all_reduce_gradients
        compressor.grad_zero = self.zero_gradients
        try:
            from . import ops
            ops.noise_source.configure(self.rank, self.world_size)
        except Exception:  # pragma: no cover - CPU-only unit tests of the bucket logic
            pass
        self.broadcast_parameters()

    # ------------------------------------------------------------------ bucket + hooks
    def rebuild_bucket(self) -> None:
        """(Re-)homes every main-parameter gradient in the flat buffer and re-arms the arrival hooks."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        main = list(self.module.get_main_parameters()) + list(self.module.loss_balancer.parameters())
        self.bucket = FlatGradBucket(main, n_buckets=self.n_buckets if self.overlap else 1)
        nb = len(self.bucket.slices)
        self._arrived = [0] * nb
        self._streams = [set() for _ in range(nb)]
        self._launched = [False] * nb
        self._work = []
        backend = dist.get_backend(self.group) if dist.is_initialized() else "none"
        self._avg = backend == "nccl"   # gloo has no AVG: sum, then scale
        if self.overlap and self.world_size > 1:
            for p in self.bucket.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self.bucket.bucket_of[id(p)]
        self._arrived[b] += 1
        if p.is_cuda:
            self._streams[b].add(torch.cuda.current_stream(p.device))
        if self._arrived[b] == self.bucket.counts[b] and not self._launched[b]:
            self._launch(b)

    def _launch(self, b: int) -> None:
        sl = self.bucket.slices[b]
        if sl.is_cuda:
            cur = torch.cuda.current_stream(sl.device)
            for s in self._streams[b]:
                if s != cur:
                    cur.wait_stream(s)  # every contribution to this bucket is ordered before the collective
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._work.append((dist.all_reduce(sl, op=op, group=self.group, async_op=True), b))
        self._launched[b] = True

    # ------------------------------------------------------------------ step protocol
    def zero_gradients(self) -> None:
        if not self.bucket.check_views():
            raise RuntimeError("a gradient left the flat bucket (zero_grad(set_to_none=True) somewhere?)")
        self.bucket.zero()

    def broadcast_parameters(self) -> None:
        if self.world_size == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            if t.numel() > 0:
                dist.broadcast(t.data, src=0, group=self.group)

    def all_reduce_gradients(self) -> None:
        """Issue the buckets that have not gone out yet, wait for all of them, reset the arrival state."""
        if self.world_size == 1:
            return
        if not self.bucket.check_views():
            raise RuntimeError("a gradient left the flat bucket (zero_grad(set_to_none=True) somewhere?)")
        for b in range(len(self.bucket.slices)):
            if not self._launched[b]:
                self._launch(b)
        for work, b in self._work:
            work.wait()
            if not self._avg:
                self.bucket.slices[b].mul_(1.0 / self.world_size)
        self._work = []
        for b in range(len(self.bucket.slices)):
            self._arrived[b], self._launched[b] = 0, False
            self._streams[b].clear()

    def training_step(self, batch, batch_idx: int = 0):
        return self.module.training_step(batch, batch_idx)

    def __getattr__(self, name):
        return getattr(self.module, name)
