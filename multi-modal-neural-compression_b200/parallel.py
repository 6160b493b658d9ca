"""Data parallelism over the image batch: one process per GPU, one NCCL all-reduce of gradients per step.

The reference has no parallelism at all (its Trainer is pinned to a single device,
/root/reference/src/train.py:288-295).  Every op on the rate path is per image, parameters are
replicated, so the path shards by batch with a single exchange step in training: the sum of the main-parameter
gradients.  Eval, compress and decompress need no collective.

Gradients live in ONE flat fp32 buffer whose slices are the `.grad` views of the parameters, so a single
`all_reduce` (NCCL over NVLink/NVSwitch on a B200 box; gloo in the CPU tests) covers the whole model.  The
`quantiles` parameters are left out: their gradient comes from the data-independent auxiliary loss
(/root/reference/src/models/multi_task_compressor.py:456-462) and is identical on every rank.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Re-homes the gradients of `params` into one contiguous buffer (in reverse registration order, i.e. roughly
    the order backward produces them)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        off = 0
        for p in reversed(self.params):
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def check_views(self) -> bool:
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


class DataParallel:
    """Wraps a compressor: after `loss.backward()` the training step calls `grad_sync`, which averages the flat
    gradient bucket across ranks."""

    def __init__(self, compressor, process_group: Optional[dist.ProcessGroup] = None):
        self.module = compressor
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        main = list(compressor.get_main_parameters()) + list(compressor.loss_balancer.parameters())
        self.bucket = FlatGradBucket(main)
        compressor.grad_sync = self.all_reduce_gradients
        try:
            from . import ops
            ops.noise_source.configure(self.rank, self.world_size)
        except Exception:  # pragma: no cover - CPU-only unit tests of the bucket logic
            pass
        self.broadcast_parameters()

    def broadcast_parameters(self) -> None:
        if self.world_size == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            if t.numel() > 0:
                dist.broadcast(t.data, src=0, group=self.group)

    def all_reduce_gradients(self) -> None:
        if self.world_size == 1:
            return
        if not self.bucket.check_views():
            raise RuntimeError("a gradient left the flat bucket (zero_grad(set_to_none=True) somewhere?)")
        dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.bucket.flat.mul_(1.0 / self.world_size)

    def training_step(self, batch, batch_idx: int = 0):
        return self.module.training_step(batch, batch_idx)

    def __getattr__(self, name):
        return getattr(self.module, name)
