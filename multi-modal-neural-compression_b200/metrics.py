"""(f2) The reference's per-step metrics, `MultiTaskCompressor.average_metrics`
(/root/reference/src/models/multi_task_compressor.py:359-384): PSNR (torchmetrics `peak_signal_noise_ratio`) and
MS-SSIM (pytorch_msssim `ms_ssim`) of every task, images scaled by 255 (data_range 255); for `semantic` the argmax
class-id image against the class-id target (data_range 17).  The reference evaluates both on every training step.

Here PSNR is free: the distortion kernel already reduced sum (x_hat - x)^2 / (B C), and
PSNR = 10 log10(data_range^2 / MSE) = -10 log10(MSE of the unscaled images).  The semantic class-id image and its
squared error come from one fused pass over the logits (`mmnc_argmax_sse`).  MS-SSIM of CUDA tensors runs on
`mmnc_ssim_scale` (csrc/ssim.cu): one fused launch per scale - the five Gaussian-filtered maps, the SSIM / contrast
maps, their spatial means and the 2 x 2 pooled input of the next scale - instead of ten depth-wise convolutions and
~20 element-wise passes (21 ms -> 1 ms for a (256, 3, 256, 256) pair).  `ms_ssim_torch` is the same computation with
stock torch ops: what CPU tensors get, and what the kernel is tested against.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib, ops

MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def semantic_labels_and_sse(logits: Tensor, target: Optional[Tensor]):
    """logits (B, K, H, W) -> (argmax class ids as float (B, 1, H, W), sum (argmax - target)^2 or None)."""
    ops._need_cuda(logits, target)
    logits = ops._f32c(logits.detach())
    B, K = logits.shape[:2]
    S = logits[0, 0].numel()
    labels = torch.empty((B, 1) + tuple(logits.shape[2:]), dtype=torch.float32, device=logits.device)
    sse = None
    tptr = None
    if target is not None:
        target = ops._f32c(target.detach())
        if target.numel() != B * S:
            raise ValueError(f"target {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
        sse = torch.zeros((), dtype=torch.float32, device=logits.device)
        tptr = ops._p(target)
    _lib.check(_lib.lib().mmnc_argmax_sse(ops._p(logits), tptr, B, K, S, ops._p(labels), ops._p(sse), ops._stream()))
    return labels, sse


def psnr_from_mse(mse: Tensor, data_range: float, scale: float = 1.0) -> Tensor:
    """10 log10(data_range^2 / (scale^2 mse)): PSNR of images that were multiplied by `scale` before the comparison."""
    return 10.0 * torch.log10((data_range * data_range) / (scale * scale * mse))


def _gaussian_window(size: int = 11, sigma: float = 1.5, device=None) -> Tensor:
    coords = torch.arange(size, dtype=torch.float32, device=device) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x: Tensor, win: Tensor) -> Tensor:
    C = x.shape[1]
    k = win.numel()
    x = F.conv2d(x, win.view(1, 1, k, 1).expand(C, 1, k, 1), groups=C)
    return F.conv2d(x, win.view(1, 1, 1, k).expand(C, 1, 1, k), groups=C)


def _ssim_cs(x: Tensor, y: Tensor, win: Tensor, data_range: float):
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = _filter(x * x, win) - mu1_sq, _filter(y * y, win) - mu2_sq, _filter(x * y, win) - mu12
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)  # (B, C) each


def _check_ms_ssim_args(x: Tensor, y: Tensor):
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError(f"ms_ssim expects two (B, C, H, W) tensors of one shape, got {tuple(x.shape)} / {tuple(y.shape)}")
    if min(x.shape[-2:]) <= (11 - 1) * 2 ** 4:
        raise ValueError("image side must be larger than 160 for five scales with an 11-tap window")


_WEIGHT_CACHE: Dict[tuple, Tensor] = {}


def _combine_scales(stack: Tensor, size_average: bool) -> Tensor:
    """(5, B, C): relu'd contrast means of scales 0-3 and the SSIM mean of scale 4 -> prod_i stack_i ^ w_i, written as
    exp(sum w_i ln stack_i) (0 ^ w = exp(-inf) = 0): three plain element-wise kernels, no run-time compiled pow / prod."""
    key = (str(stack.device), stack.dtype)
    if key not in _WEIGHT_CACHE:
        _WEIGHT_CACHE[key] = torch.tensor(MS_SSIM_WEIGHTS, dtype=stack.dtype, device=stack.device).view(-1, 1, 1)
    val = torch.exp((torch.log(stack) * _WEIGHT_CACHE[key]).sum(0))
    return val.mean() if size_average else val.mean(1)


def ms_ssim(x: Tensor, y: Tensor, data_range: float = 255.0, size_average: bool = True, scale: float = 1.0) -> Tensor:
    """Multi-scale SSIM as pytorch_msssim computes it: 11-tap Gaussian (sigma 1.5), 5 scales with 2x2 average pooling
    in between, contrast-structure terms of the first four scales and the full SSIM of the last, ReLU'd, combined
    with the published exponents, averaged over channels (and the batch).  `scale` multiplies both images first (the
    reference compares x * 255 with data_range 255); on CUDA tensors that costs nothing (applied on load)."""
    _check_ms_ssim_args(x, y)
    if not (x.is_cuda and y.is_cuda):
        return ms_ssim_torch(x * scale if scale != 1.0 else x, y * scale if scale != 1.0 else y, data_range, size_average)
    L = _lib.lib()
    x, y = ops._f32c(x.detach()), ops._f32c(y.detach())
    B, C, H, W = x.shape
    planes = B * C
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    means = torch.empty(5, 2, planes, dtype=torch.float32, device=x.device)  # [scale][ssim | cs][plane]
    if H % 16 == 0 and W % 16 == 0:  # every pooling is an exact 2 x 2: all five scales in ONE call (ten launches)
        ws = torch.empty(int(L.mmnc_ms_ssim_workspace_floats(planes, H, W)), dtype=torch.float32, device=x.device)
        _lib.check(L.mmnc_ms_ssim(ops._p(x), ops._p(y), planes, H, W, scale, c1, c2, 1.5, ops._p(ws), ops._p(means),
                                  ops._stream()))
        stack = torch.cat([means[:4, 1], means[4:, 0]], dim=0).view(5, B, C)
        return _combine_scales(torch.relu(stack), size_average)
    ws = torch.empty(int(L.mmnc_ssim_workspace_floats(planes, H, W)), dtype=torch.float32, device=x.device)
    for i in range(5):
        H, W = x.shape[-2:]
        fused_pool = i < 4 and H % 2 == 0 and W % 2 == 0
        px = py = None
        if fused_pool:
            px = torch.empty(B, C, H // 2, W // 2, dtype=torch.float32, device=x.device)
            py = torch.empty_like(px)
        _lib.check(L.mmnc_ssim_scale(ops._p(x), ops._p(y), planes, H, W, scale if i == 0 else 1.0, c1, c2, 1.5, ops._p(ws),
                                     ops._p(means[i, 0]), ops._p(means[i, 1]), ops._p(px) if fused_pool else None,
                                     ops._p(py) if fused_pool else None, ops._stream()))
        if i < 4:
            if fused_pool:
                x, y = px, py
            else:  # odd sides: pytorch_msssim pads the pooling with zeros (count_include_pad), rare enough for torch ops
                pad = [s % 2 for s in x.shape[2:]]
                s0 = scale if i == 0 else 1.0
                x, y = F.avg_pool2d(x * s0, 2, padding=pad), F.avg_pool2d(y * s0, 2, padding=pad)
    stack = torch.cat([means[:4, 1], means[4:, 0]], dim=0).view(5, B, C)
    return _combine_scales(torch.relu(stack), size_average)


def ms_ssim_torch(x: Tensor, y: Tensor, data_range: float = 255.0, size_average: bool = True) -> Tensor:
    """The same computation with stock torch ops (any device)."""
    _check_ms_ssim_args(x, y)
    win = _gaussian_window(device=x.device)
    weights = torch.tensor(MS_SSIM_WEIGHTS, dtype=x.dtype, device=x.device)
    mcs = []
    for i in range(5):
        ssim_c, cs = _ssim_cs(x, y, win, data_range)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x, y = F.avg_pool2d(x, 2, padding=pad), F.avg_pool2d(y, 2, padding=pad)
    stack = torch.stack(mcs + [torch.relu(ssim_c)], dim=0)  # (5, B, C)
    val = torch.prod(stack ** weights.view(-1, 1, 1), dim=0)
    return val.mean() if size_average else val.mean(1)


def average_metrics(tasks, x: Dict[str, Tensor], x_hats: Dict[str, Tensor], log_dir: str,
                    task_losses: Optional[Dict[str, Tensor]] = None, with_ms_ssim: bool = True) -> Dict[str, Tensor]:
    """-> {f"{log_dir}/{task}/psnr": ..., f"{log_dir}/{task}/ms-ssim": ...} like mtc.py:359-384.

    `task_losses[task]` (optional) = the distortion term sum (x_hat - x)^2 / (B C) already computed for the RD loss:
    PSNR then costs nothing; otherwise it is computed with the same distortion kernel."""
    logs: Dict[str, Tensor] = {}
    with torch.no_grad():
        for task in tasks:
            pred, tgt = x_hats[task].detach(), x[task]
            if task == "semantic":
                labels, sse = semantic_labels_and_sse(pred, tgt)
                mse = sse / tgt.numel()
                logs[f"{log_dir}/{task}/psnr"] = psnr_from_mse(mse, 17.0)
                if with_ms_ssim:
                    logs[f"{log_dir}/{task}/ms-ssim"] = ms_ssim(labels, tgt.float(), data_range=17.0)
                continue
            hw = pred.shape[-2] * pred.shape[-1]
            if task_losses is not None and task in task_losses:
                mse = task_losses[task].detach() / hw
            else:
                mse = ops.distortion(pred, tgt, "mse") / hw
            logs[f"{log_dir}/{task}/psnr"] = psnr_from_mse(mse, 255.0, 255.0)
            if with_ms_ssim:
                logs[f"{log_dir}/{task}/ms-ssim"] = ms_ssim(pred * 255.0, tgt * 255.0, data_range=255.0)
    return logs
