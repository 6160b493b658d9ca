"""(f4) GPU-side input pipeline: batches of raw decoded pixels -> the dict of float NCHW tensors the compressors take.

Reference: `CLEVRDataset.__getitem__` + `get_transform` convert every sample on the host (PIL -> numpy -> fp32 CHW)
inside DataLoader worker processes, the default collate stacks them and Lightning ships fp32 to the device
(/root/reference/src/datasets/clevr.py:48-83, /root/reference/src/datasets/transforms.py:39-131,
/root/reference/src/train.py:161-200 - no pin_memory, no prefetch).  Here the workers only decode the PNGs; a batch
crosses PCIe as uint8 / uint16 (29 MB instead of 117 MB for 64 images x 3 tasks) from pinned memory on a copy stream,
one step ahead of the consumer, and three kernels (csrc/prep.cu) produce exactly the tensors the reference's transforms
produce.  PNG decoding and resizing stay on the host (out of scope; CLEVR-Taskonomy is stored at 256 x 256).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops

SEM1_CLASSES = (0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15, 16, 17, 255)  # src/datasets/clevr.py:13

# task -> (raw dtype, output channels, kind)
RAW_FORMAT = {
    "rgb": (np.uint8, 3, "u8"), "normal": (np.uint8, 3, "u8"), "mono": (np.uint8, 1, "u8"),
    "depth_euclidean": (np.uint16, 1, "u16"), "semantic": (np.uint8, 1, "labels"),
}


def semantic_lut(device) -> Tensor:
    """256-entry table of the reference's replacement loop `for i, class_ in enumerate(SEM1_CLASSES): x[x == class_] = i`
    (clevr.py:76-77), INCLUDING its sequential aliasing: the loop rewrites the tensor in place, so a value that has
    already been mapped can be matched again by a later class (raw 10 -> 8, but raw 8 is no class and stays 8)."""
    lut = np.arange(256, dtype=np.int64)
    for i, cls in enumerate(SEM1_CLASSES):
        lut[lut == cls] = i
    return torch.from_numpy(lut.astype(np.float32)).to(device)


def convert_raw(task: str, raw: Tensor, lut: Optional[Tensor] = None) -> Tensor:
    """One batch of raw pixels already on the device -> float (B, C, H, W), as the reference's transform would give.
    raw: uint8 (B, H, W, Cs) for rgb / normal / mono / semantic (or (B, H, W) for single-channel data), uint16
    (B, H, W) viewed as int16 storage for depth."""
    ops._need_cuda(raw)
    dtype, cd, kind = RAW_FORMAT[task]
    raw = raw.contiguous()
    B, H, W = raw.shape[:3]
    L, st = _lib.lib(), ops._stream()
    out = torch.empty((B, cd, H, W), dtype=torch.float32, device=raw.device)
    if kind == "u16":
        if raw.dtype not in (torch.int16, torch.uint16) or raw.dim() != 3:
            raise TypeError(f"{task}: expected a 16-bit (B, H, W) tensor, got {raw.dtype} {tuple(raw.shape)}")
        _lib.check(L.mmnc_prep_u16_to_f32(ops._p(raw), raw.numel(), float(2 ** 15 - 1), ops._p(out), st))
        return out
    if raw.dtype != torch.uint8:
        raise TypeError(f"{task}: expected uint8 pixels, got {raw.dtype}")
    cs = 1 if raw.dim() == 3 else raw.shape[3]
    if kind == "u8":
        if cs < cd:
            raise ValueError(f"{task}: {cs} source channels, {cd} needed")
        _lib.check(L.mmnc_prep_u8_hwc_to_f32_chw(ops._p(raw), B, H * W, cs, cd, 255.0, ops._p(out), st))
        return out
    if cs < 2:
        raise ValueError("semantic: the class image is channel 1 of a multi-channel PNG (clevr.py:68-73)")
    lut = semantic_lut(raw.device) if lut is None else lut
    _lib.check(L.mmnc_prep_labels(ops._p(raw), B * H * W, cs, 1, ops._p(lut), ops._p(out), st))
    return out


class GpuBatchLoader:
    """Wraps an iterable of {task: numpy array of raw pixels, batch first} and yields {task: float CUDA tensor}.

    Every batch is copied into pinned staging memory, sent on a private copy stream and converted there; the NEXT
    batch is in flight while the caller works on the current one (double buffering).  The consumer's stream waits on
    the copy stream before it touches a batch."""

    def __init__(self, source: Iterable[Dict[str, np.ndarray]], tasks: Sequence[str], device="cuda"):
        self.source, self.tasks = source, tuple(tasks)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = [dict(), dict()]
        self._slot_done = [None, None]  # CUDA events: the copies out of each pinned slot have completed
        self._lut = None
        self.h2d_bytes_per_batch = 0

    def _stage(self, slot: int, raw: Dict[str, np.ndarray]) -> Dict[str, Tensor]:
        out, nbytes = {}, 0
        for t in self.tasks:
            a = np.ascontiguousarray(raw[t])
            if a.dtype == np.uint16:
                a = a.view(np.int16)  # same bits; torch's uint16 support is partial
            buf = self._pinned[slot].get(t)
            if buf is None or buf.shape != a.shape or buf.dtype != torch.from_numpy(a).dtype:
                buf = self._pinned[slot][t] = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
            buf.numpy()[...] = a
            nbytes += a.nbytes
            out[t] = buf
        self.h2d_bytes_per_batch = nbytes
        return out

    def _launch(self, slot: int, raw) -> Dict[str, Tensor]:
        if self._slot_done[slot] is not None:
            self._slot_done[slot].synchronize()  # never overwrite pinned memory a copy may still be reading
        staged = self._stage(slot, raw)
        with torch.cuda.stream(self.stream):
            if self._lut is None and "semantic" in self.tasks:
                self._lut = semantic_lut(self.device)
            out = {t: convert_raw(t, staged[t].to(self.device, non_blocking=True), self._lut) for t in self.tasks}
            self._slot_done[slot] = torch.cuda.Event()
            self._slot_done[slot].record(self.stream)
        return out

    def __iter__(self) -> Iterator[Dict[str, Tensor]]:
        it = iter(self.source)
        slot = 0
        try:
            pending = self._launch(slot, next(it))
        except StopIteration:
            return
        done = torch.cuda.Event()
        done.record(self.stream)
        while pending is not None:
            cur, cur_done = pending, done
            slot ^= 1
            try:
                pending = self._launch(slot, next(it))
                done = torch.cuda.Event()
                done.record(self.stream)
            except StopIteration:
                pending = None
            main = torch.cuda.current_stream(self.device)
            main.wait_event(cur_done)
            for v in cur.values():
                v.record_stream(main)
            yield cur
