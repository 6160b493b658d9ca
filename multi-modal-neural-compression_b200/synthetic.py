"""Synthetic CLEVR-shaped multi-task batches (value ranges from the reference's loaders:
/root/reference/src/datasets/clevr.py:13,74-77 and /root/reference/src/plots.ipynb:443; seed 21 as in
/root/reference/src/train.py:204).  There is no dataset on the box, so every benchmark uses these."""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from .compressors import task_parameters


def synthetic_batch(tasks: Sequence[str], batch_size: int, size: int = 256, seed: int = 21, device="cpu",
                    pin_memory: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for t in tasks:
        c = task_parameters[t]["in_channels"]
        if t == "semantic":
            v = torch.randint(0, 17, (batch_size, c, size, size), generator=g).float()
        elif t == "depth_euclidean":
            v = torch.rand(batch_size, c, size, size, generator=g) * 4.1
        else:
            v = torch.rand(batch_size, c, size, size, generator=g)
        if pin_memory:
            v = v.pin_memory()
        out[t] = v.to(device)
    return out
