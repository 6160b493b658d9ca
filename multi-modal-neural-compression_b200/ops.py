"""Autograd-aware Python entry points over the C ABI (include/mmnc_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every arithmetic step of the rate path
runs in libmmnc_b200.so.  All ops require CUDA tensors and raise otherwise — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

QUANT_DEQUANTIZE, QUANT_NOISE_PHILOX, QUANT_NOISE_GIVEN, QUANT_IDENTITY, QUANT_NOISE_PHILOX_DEV = 0, 1, 2, 3, 4
MEANS_NONE, MEANS_PER_CHANNEL, MEANS_FULL = 0, 1, 2
EB_FORM = {"sign": 0, "plain": 1}
GDN_PRECISION = {"fp32": 0, "tf32": 1, "3xtf32": 2, "auto": 3}
EB_NP = 58


def _p(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "mmnc_b200: expected CUDA tensors — the rate path has no CPU / PyTorch fallback "
                f"(got a tensor on {t.device})")


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"mmnc_b200: float32 expected, got {t.dtype}")
    return t.contiguous()


def _bcs(t: Tensor) -> Tuple[int, int, int]:
    if t.dim() < 2:
        raise ValueError("expected a tensor with at least 2 dimensions (B, C, ...)")
    B, C = t.shape[0], t.shape[1]
    S = 1
    for d in t.shape[2:]:
        S *= d
    return B, C, S


class _NoiseSeed:
    """Philox (seed, offset) source.  `offset` lets data-parallel ranks draw from disjoint parts of one stream
    so that results do not depend on the world size."""

    def __init__(self):
        self.calls = 0
        self.rank, self.world_size = 0, 1
        self.state: Optional[Tensor] = None
        self.slot = 0

    def configure(self, rank: int = 0, world_size: int = 1) -> None:
        self.rank, self.world_size = int(rank), int(world_size)

    # ---- device-resident (seed, offset): lets a captured CUDA graph draw fresh noise on every replay
    def enable_device_state(self, device, seed: int = 21) -> None:
        self.state = torch.tensor([int(seed), 0], dtype=torch.int64, device=device)
        self.slot = 0

    def disable_device_state(self) -> None:
        self.state = None

    def step(self) -> None:
        """Advance the device-side stream (call once per training step, inside the captured region)."""
        if self.state is not None:
            self.state[0:1].add_(1)
            self.slot = 0

    def next_device(self, numel: int) -> Tuple[Tensor, int]:
        """-> (state tensor, host offset): every call inside a step gets its own 2^40-element slot."""
        self.slot += 1
        return self.state, self.slot * (1 << 40) + self.rank * int(numel)

    def next(self, numel: int = 0) -> Tuple[int, int]:
        # The seed follows torch's CPU generator, so torch.manual_seed() makes runs reproducible and ranks seeded
        # alike draw the same seed; rank r then reads elements [r * numel, (r + 1) * numel) of that stream.
        self.calls += 1
        seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())
        return seed, self.rank * int(numel)


noise_source = _NoiseSeed()


# ------------------------------------------------------------------------------------------------ quantise (a2)
def quantize_noise(x: Tensor, noise: Optional[Tensor] = None, seed: Optional[int] = None, offset: int = 0,
                   state: Optional[Tensor] = None) -> Tensor:
    _need_cuda(x, noise)
    x = _f32c(x)
    out = torch.empty_like(x)
    if state is not None:  # device-resident (seed, offset)
        mode, noise, seed = QUANT_NOISE_PHILOX_DEV, state, 0
    elif noise is not None:
        mode, noise = QUANT_NOISE_GIVEN, _f32c(noise.expand_as(x))
        seed = 0
    else:
        mode = QUANT_NOISE_PHILOX
        if seed is None:
            seed, offset = noise_source.next(x.numel())
    _lib.check(_lib.lib().mmnc_quantize_noise(_p(x), x.numel(), mode, _p(noise), seed, offset, _p(out), _stream()))
    return out


def _means_arg(x: Tensor, means: Optional[Tensor]):
    """Classifies `means` as none / per-channel / full for the kernels; anything else is expanded to full."""
    if means is None:
        return None, MEANS_NONE
    _need_cuda(means)
    C = x.shape[1]
    if means.numel() == C and means.dim() == x.dim() and means.shape[1] == C:
        return _f32c(means).reshape(-1), MEANS_PER_CHANNEL
    return _f32c(means.expand_as(x)), MEANS_FULL


def quantize_dequantize(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    _need_cuda(x)
    x = _f32c(x)
    B, C, S = _bcs(x)
    m, mode = _means_arg(x, means)
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mmnc_quantize_dequantize(_p(x), B, C, S, _p(m), mode, _p(out), _stream()))
    return out


def quantize_symbols(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    _need_cuda(x)
    x = _f32c(x)
    B, C, S = _bcs(x)
    m, mode = _means_arg(x, means)
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().mmnc_quantize_symbols(_p(x), B, C, S, _p(m), mode, _p(out), _stream()))
    return out


def dequantize_symbols(symbols: Tensor, means: Optional[Tensor] = None) -> Tensor:
    _need_cuda(symbols)
    symbols = symbols.contiguous()
    if symbols.dtype != torch.int32:
        symbols = symbols.int()
    B, C, S = _bcs(symbols)
    m, mode = _means_arg(symbols, means)
    out = torch.empty(symbols.shape, dtype=torch.float32, device=symbols.device)
    _lib.check(_lib.lib().mmnc_dequantize_symbols(_p(symbols), B, C, S, _p(m), mode, _p(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ EB (a3, a4)
class _EntropyBottleneckFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, packed, medians, mode, noise, seed, offset, bound, form):
        _need_cuda(x, packed, medians, noise)
        x, packed, medians = _f32c(x), _f32c(packed), _f32c(medians)
        noise = noise if mode == QUANT_NOISE_PHILOX_DEV else _f32c(noise)
        B, C, S = _bcs(x)
        if packed.shape != (C, EB_NP):
            raise ValueError(f"packed EB parameters must be ({C}, {EB_NP}), got {tuple(packed.shape)}")
        out, lik = torch.empty_like(x), torch.empty_like(x)
        lnsum = torch.zeros(C, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mmnc_eb_forward(_p(x), B, C, S, _p(packed), _p(medians.reshape(-1)), mode, _p(noise),
                                              seed, offset, bound, form, _p(out), _p(lik), _p(lnsum), _stream()))
        ctx.save_for_backward(out, packed)
        ctx.dims, ctx.cfg = (B, C, S), (mode, bound, form)
        ctx.set_materialize_grads(False)
        return out, lik, lnsum

    @staticmethod
    def backward(ctx, g_out, g_lik, g_lnsum):
        out, packed = ctx.saved_tensors
        B, C, S = ctx.dims
        mode, bound, form = ctx.cfg
        g_out, g_lik, g_lnsum = _f32c(g_out), _f32c(g_lik), _f32c(g_lnsum)
        g_x = torch.empty_like(out)
        g_packed = torch.zeros_like(packed)
        _lib.check(_lib.lib().mmnc_eb_backward(_p(out), B, C, S, _p(packed), _p(g_out), _p(g_lik), _p(g_lnsum), bound,
                                               form, _p(g_x), _p(g_packed), _stream()))
        if mode == QUANT_DEQUANTIZE:
            g_x.zero_()  # round() has zero gradient
        return g_x, g_packed, None, None, None, None, None, None, None


def entropy_bottleneck_forward(x: Tensor, packed: Tensor, medians: Tensor, training: bool, bound: float,
                               form: str = "sign", noise: Optional[Tensor] = None, seed: Optional[int] = None,
                               offset: int = 0, quantized: bool = False):
    """-> (outputs, likelihood, lnsum[C]) where lnsum[c] = sum over (b, spatial) of ln(likelihood)."""
    if quantized:
        mode = QUANT_IDENTITY
    elif not training:
        mode = QUANT_DEQUANTIZE
    elif noise is not None:
        mode, noise = QUANT_NOISE_GIVEN, noise.expand_as(x)
    elif seed is None and noise_source.state is not None:
        mode = QUANT_NOISE_PHILOX_DEV
        noise, offset = noise_source.next_device(x.numel())
    else:
        mode = QUANT_NOISE_PHILOX
        if seed is None:
            seed, offset = noise_source.next(x.numel())
    return _EntropyBottleneckFn.apply(x, packed, medians.detach(), mode, noise, seed or 0, offset, float(bound),
                                      EB_FORM[form])


def eb_logits(v: Tensor, packed: Tensor) -> Tensor:
    """_logits_cumulative(v, stop_gradient=True) for v of shape (C, 1, L) or (C, L)."""
    _need_cuda(v, packed)
    v2 = _f32c(v.detach()).reshape(v.shape[0], -1)
    out = torch.empty_like(v2)
    _lib.check(_lib.lib().mmnc_eb_logits(_p(v2), v2.shape[0], v2.shape[1], _p(_f32c(packed.detach())), _p(out), _stream()))
    return out.reshape(v.shape)


class _EBAuxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, quantiles, packed, target):
        _need_cuda(quantiles, packed, target)
        q = _f32c(quantiles)
        C = q.shape[0]
        loss = torch.zeros((), dtype=torch.float32, device=q.device)
        gq = torch.empty_like(q)
        _lib.check(_lib.lib().mmnc_eb_aux_loss(_p(q), C, _p(_f32c(packed)), _p(_f32c(target)), _p(loss), _p(gq), _stream()))
        ctx.save_for_backward(gq)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gq,) = ctx.saved_tensors
        return gq * g, None, None


def eb_aux_loss(quantiles: Tensor, packed: Tensor, target: Tensor) -> Tensor:
    return _EBAuxLossFn.apply(quantiles, packed.detach(), target)


# ------------------------------------------------------------------------------------------------ GC (a5)
class _GaussianConditionalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, scales, means, mode, noise, seed, offset, scale_bound, lik_bound):
        _need_cuda(y, scales, means, noise)
        y, scales, means = _f32c(y), _f32c(scales), _f32c(means)
        noise = noise if mode == QUANT_NOISE_PHILOX_DEV else _f32c(noise)
        B, C, Sy = _bcs(y)
        B2, C2, Ss = _bcs(scales)
        if (B, C) != (B2, C2) or not (Sy == Ss or Sy == 1):
            raise ValueError(f"incompatible shapes {tuple(y.shape)} / {tuple(scales.shape)}")
        y_hat = torch.empty_like(y)
        lik = torch.empty_like(scales)
        lnsum = torch.zeros(C, dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib().mmnc_gc_forward(_p(y), _p(scales), _p(means), B, C, Sy, Ss, mode, _p(noise), seed, offset,
                                              scale_bound, lik_bound, _p(y_hat), _p(lik), _p(lnsum), _stream()))
        ctx.save_for_backward(y_hat, scales, means)
        ctx.dims, ctx.cfg = (B, C, Sy, Ss), (mode, scale_bound, lik_bound)
        ctx.set_materialize_grads(False)
        return y_hat, lik, lnsum

    @staticmethod
    def backward(ctx, g_yhat, g_lik, g_lnsum):
        y_hat, scales, means = ctx.saved_tensors
        B, C, Sy, Ss = ctx.dims
        mode, scale_bound, lik_bound = ctx.cfg
        g_yhat, g_lik, g_lnsum = _f32c(g_yhat), _f32c(g_lik), _f32c(g_lnsum)
        g_y, g_scales = torch.empty_like(y_hat), torch.empty_like(scales)
        _lib.check(_lib.lib().mmnc_gc_backward(_p(y_hat), _p(scales), _p(means), B, C, Sy, Ss, _p(g_yhat), _p(g_lik),
                                               _p(g_lnsum), scale_bound, lik_bound, _p(g_y), _p(g_scales), _stream()))
        if mode == QUANT_DEQUANTIZE:
            g_y.zero_()
        return g_y, g_scales, None, None, None, None, None, None, None


def gaussian_conditional_forward(y: Tensor, scales: Tensor, means: Optional[Tensor], training: bool,
                                 scale_bound: float, lik_bound: float, noise: Optional[Tensor] = None,
                                 seed: Optional[int] = None, offset: int = 0):
    """-> (outputs, likelihood, lnsum[C]).  Handles torch-style broadcasting between y and scales."""
    same = y.shape == scales.shape
    y_is_point = (y.dim() == scales.dim() and y.shape[:2] == scales.shape[:2]
                  and all(d == 1 for d in y.shape[2:]))
    if means is not None:
        if means.requires_grad:
            raise NotImplementedError("mmnc_b200: gradients w.r.t. `means` are not implemented (ScaleHyperprior has none)")
        means = means.expand_as(y)
    if not training:
        mode = QUANT_DEQUANTIZE
    elif noise is not None:
        mode, noise = QUANT_NOISE_GIVEN, noise.expand_as(y)
    elif seed is None and noise_source.state is not None:
        mode = QUANT_NOISE_PHILOX_DEV
        noise, offset = noise_source.next_device(y.numel())
    else:
        mode = QUANT_NOISE_PHILOX
        if seed is None:
            seed, offset = noise_source.next(y.numel())
    if same or y_is_point:
        return _GaussianConditionalFn.apply(y, scales, means, mode, noise, seed or 0, offset, float(scale_bound),
                                            float(lik_bound))
    # general broadcasting: quantise y in its own shape, then evaluate on the materialised broadcast
    if mode == QUANT_DEQUANTIZE:
        y_hat = _StraightRound.apply(y, means)
    elif mode == QUANT_NOISE_PHILOX_DEV:
        y_hat = _AddNoise.apply(y, None, None, offset, noise)
    else:
        y_hat = _AddNoise.apply(y, noise, seed, offset)
    yb, sb = torch.broadcast_tensors(y_hat, scales)
    mb = None if means is None else means.expand_as(yb)
    _, lik, lnsum = _GaussianConditionalFn.apply(yb.contiguous(), sb.contiguous(), mb, QUANT_IDENTITY, None, 0, 0,
                                                 float(scale_bound), float(lik_bound))
    return y_hat, lik, lnsum


class _AddNoise(torch.autograd.Function):
    """x + U(-1/2, 1/2) (or + given noise) with d/dx = 1."""

    @staticmethod
    def forward(ctx, x, noise, seed, offset, state=None):
        return quantize_noise(x, noise=noise, seed=seed, offset=offset, state=state)

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None


class _StraightRound(torch.autograd.Function):
    """round(x - m) + m with torch.round's zero gradient."""

    @staticmethod
    def forward(ctx, x, means):
        return quantize_dequantize(x, means)

    @staticmethod
    def backward(ctx, g):
        return torch.zeros_like(g), None


# ------------------------------------------------------------------------------------------------ ln-sum (a6)
class _LnSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lik):
        _need_cuda(lik)
        lik = _f32c(lik)
        B, C, S = _bcs(lik)
        out = torch.zeros(C, dtype=torch.float32, device=lik.device)
        _lib.check(_lib.lib().mmnc_lnsum_forward(_p(lik), B, C, S, _p(out), _stream()))
        ctx.save_for_backward(lik)
        ctx.dims = (B, C, S)
        return out

    @staticmethod
    def backward(ctx, g):
        (lik,) = ctx.saved_tensors
        B, C, S = ctx.dims
        g_lik = torch.empty_like(lik)
        _lib.check(_lib.lib().mmnc_lnsum_backward(_p(lik), B, C, S, _p(_f32c(g)), _p(g_lik), _stream()))
        return g_lik


def channel_log_likelihood_sums(lik: Tensor) -> Tensor:
    """(B, C, ...) likelihoods -> (C,) sums of ln(lik) over batch and space."""
    return _LnSumFn.apply(lik)


# ------------------------------------------------------------------------------------------------ distortion (a7)
class _DistortionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_hat, x, kind, scale):
        _need_cuda(x_hat, x)
        if x_hat.shape != x.shape:
            raise ValueError(f"shape mismatch {tuple(x_hat.shape)} vs {tuple(x.shape)}")
        if (x_hat.stride() == x.stride() and x_hat.dtype == x.dtype == torch.float32 and _is_channels_last(x_hat)):
            a, b = x_hat, x  # an element-wise sum does not care about the order: both channels-last, no conversion
        else:
            a, b = _f32c(x_hat), _f32c(x)
        out = torch.zeros((), dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib().mmnc_distortion_forward(_p(a), _p(b), a.numel(), kind, scale, _p(out), _stream()))
        ctx.save_for_backward(a, b)
        ctx.cfg = (kind, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        kind, scale = ctx.cfg
        g_a = torch.empty_like(a)
        _lib.check(_lib.lib().mmnc_distortion_backward(_p(a), _p(b), a.numel(), kind, scale, _p(_f32c(g)), _p(g_a),
                                                       _stream()))
        return g_a, None, None, None


def distortion(x_hat: Tensor, x: Tensor, kind: str) -> Tensor:
    """sum over all elements of d(x_hat, x) / (B * C): the reference's `mse` / `l1` terms (mtc.py:236-243)."""
    scale = 1.0 / (x.shape[0] * x.shape[1])
    return _DistortionFn.apply(x_hat, x, {"mse": 0, "l1": 1}[kind], scale)


# ------------------------------------------------------------------------------------------------ RD epilogue
class _RDEpilogueFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lnsum_y, lnsum_z, task_losses, log_vars, group_of_channel, group_inv_pixels, group_weight,
                z_inv_pixels, z_weight, lmbda):
        _need_cuda(lnsum_y, lnsum_z, task_losses, log_vars, group_of_channel)
        lnsum_y, lnsum_z, task_losses = _f32c(lnsum_y), _f32c(lnsum_z), _f32c(task_losses)
        log_vars = _f32c(log_vars)
        M, N, T = lnsum_y.numel(), lnsum_z.numel(), task_losses.numel()
        G = group_inv_pixels.numel()
        dev = lnsum_y.device
        scalars = torch.zeros(4 + G + T, dtype=torch.float32, device=dev)
        g_y, g_z = torch.empty_like(lnsum_y), torch.empty_like(lnsum_z)
        g_t = torch.empty_like(task_losses)
        g_lv = torch.empty_like(log_vars) if log_vars is not None else None
        _lib.check(_lib.lib().mmnc_rd_epilogue(_p(lnsum_y), M, _p(lnsum_z), N, _p(group_of_channel), G,
                                               _p(group_inv_pixels), _p(group_weight), z_inv_pixels, z_weight,
                                               _p(task_losses), T, _p(log_vars), lmbda, _p(scalars), _p(g_y), _p(g_z),
                                               _p(g_t), _p(g_lv), _stream()))
        ctx.save_for_backward(g_y, g_z, g_t, g_lv)
        ctx.mark_non_differentiable(scalars)
        return scalars[0].clone(), scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        g_y, g_z, g_t, g_lv = ctx.saved_tensors
        return (g_y * g_loss, g_z * g_loss, g_t * g_loss, None if g_lv is None else g_lv * g_loss,
                None, None, None, None, None, None)


def rd_epilogue(lnsum_y, lnsum_z, task_losses, log_vars, group_of_channel, group_inv_pixels, group_weight,
                z_inv_pixels: float, z_weight: float, lmbda: float):
    """-> (loss, scalars) with scalars = [loss, rec, comp, z_bpp, group bpp..., weighted task losses...]."""
    return _RDEpilogueFn.apply(lnsum_y, lnsum_z, task_losses, log_vars, group_of_channel, group_inv_pixels,
                               group_weight, float(z_inv_pixels), float(z_weight), float(lmbda))


# ------------------------------------------------------------------------------------------------ GDN (a8)
class _NonNegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, bound, pedestal):
        _need_cuda(p)
        p = _f32c(p)
        out = torch.empty_like(p)
        _lib.check(_lib.lib().mmnc_nonneg_reparam_forward(_p(p), p.numel(), bound, pedestal, _p(out), _stream()))
        ctx.save_for_backward(p)
        ctx.bound = bound
        return out

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        g_p = torch.empty_like(p)
        _lib.check(_lib.lib().mmnc_nonneg_reparam_backward(_p(p), _p(_f32c(g)), p.numel(), ctx.bound, _p(g_p), _stream()))
        return g_p, None, None


def nonneg_reparam(p: Tensor, bound: float, pedestal: float) -> Tensor:
    return _NonNegFn.apply(p, float(bound), float(pedestal))


class _GDNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, beta, gamma, inverse, precision):
        _need_cuda(x, beta, gamma)
        x, beta, gamma = _f32c(x), _f32c(beta), _f32c(gamma)
        B, C, HW = _bcs(x)
        if beta.numel() != C or gamma.shape != (C, C):
            raise ValueError(f"GDN parameters do not match {C} channels")
        y = torch.empty_like(x)
        ws, nbytes = _gdn_forward_workspace(x, B, C, HW, precision)
        _lib.check(_lib.lib().mmnc_gdn_forward_ws(_p(x), B, C, HW, _p(beta), _p(gamma), int(inverse), precision, _p(y),
                                                  _p(ws) if ws is not None else None, nbytes, _stream()))
        ctx.save_for_backward(x, beta, gamma)
        ctx.cfg = (B, C, HW, int(inverse), precision)
        return y

    @staticmethod
    def backward(ctx, g):
        x, beta, gamma = ctx.saved_tensors
        B, C, HW, inverse, precision = ctx.cfg
        g = _f32c(g)
        dx = torch.empty_like(x)
        dbeta, dgamma = torch.empty_like(beta), torch.empty_like(gamma)
        nbytes = int(_lib.lib().mmnc_gdn_backward_workspace_bytes(B, C, HW, precision))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(_lib.lib().mmnc_gdn_backward(_p(x), _p(g), B, C, HW, _p(beta), _p(gamma), inverse, precision, _p(dx),
                                                _p(dbeta), _p(dgamma), _p(ws), nbytes, _stream()))
        return dx, dbeta, dgamma, None, None


def _gdn_forward_workspace(x: Tensor, B: int, C: int, HW: int, precision: int):
    """Scratch of the wide-layer forward (129 .. 256 channels: gamma is packed there and streamed); None when not needed."""
    nbytes = int(_lib.lib().mmnc_gdn_forward_workspace_bytes(B, C, HW, precision))
    if nbytes == 0:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=x.device), nbytes


def _is_channels_last(x: Tensor) -> bool:
    """Stored NHWC and NOT also NCHW-contiguous (C = 1 or 1 x 1 images are the same bytes either way)."""
    return x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)


class _GDNRawFn(torch.autograd.Function):
    """GDN on the raw `beta` / `gamma` parameters: the re-parametrisation and its LowerBound gradient are fused
    into the contraction kernels (1 launch forward, 2 backward).

    Memory format: a channels-last (NHWC) input is processed in place where a channels-last kernel exists
    (`mmnc_gdn_nhwc_supported`) and the output keeps the input's format either way, so a model that runs its
    convolutions channels-last never converts around a GDN site of the large layers."""

    @staticmethod
    def forward(ctx, x, beta, gamma, beta_bound, gamma_bound, pedestal, inverse, precision):
        _need_cuda(x, beta, gamma)
        if x.dtype != torch.float32:
            raise TypeError(f"mmnc_b200: float32 expected, got {x.dtype}")
        beta, gamma = _f32c(beta), _f32c(gamma)
        B, C, HW = _bcs(x)
        if beta.numel() != C or gamma.shape != (C, C):
            raise ValueError(f"GDN parameters do not match {C} channels")
        L = _lib.lib()
        cl = _is_channels_last(x)
        native = cl and bool(L.mmnc_gdn_nhwc_supported(B, C, HW, precision, 0))
        if native:
            y = torch.empty_like(x)  # preserves the channels-last strides
            _lib.check(L.mmnc_gdn_forward_raw_nhwc(_p(x), B, C, HW, _p(beta), _p(gamma), beta_bound, gamma_bound,
                                                   pedestal, int(inverse), precision, _p(y), _stream()))
        else:
            x = x.contiguous()
            y = torch.empty_like(x)
            ws, nbytes = _gdn_forward_workspace(x, B, C, HW, precision)
            _lib.check(L.mmnc_gdn_forward_raw_ws(_p(x), B, C, HW, _p(beta), _p(gamma), beta_bound, gamma_bound, pedestal,
                                                 int(inverse), precision, _p(y), _p(ws) if ws is not None else None,
                                                 nbytes, _stream()))
            if cl:
                y = y.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, beta, gamma)
        ctx.cfg = (B, C, HW, int(inverse), precision, beta_bound, gamma_bound, pedestal, cl, native)
        return y

    @staticmethod
    def backward(ctx, g):
        x, beta, gamma = ctx.saved_tensors
        B, C, HW, inverse, precision, bb, gb, ped, cl, native = ctx.cfg
        L = _lib.lib()
        dbeta, dgamma = torch.empty_like(beta), torch.empty_like(gamma)
        nbytes = int(L.mmnc_gdn_backward_workspace_bytes(B, C, HW, precision))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        if native and bool(L.mmnc_gdn_nhwc_supported(B, C, HW, precision, 1)):
            g = g.contiguous(memory_format=torch.channels_last)
            dx = torch.empty_like(x)
            _lib.check(L.mmnc_gdn_backward_raw_nhwc(_p(x), _p(g), B, C, HW, _p(beta), _p(gamma), bb, gb, ped, inverse,
                                                    precision, _p(dx), _p(dbeta), _p(dgamma), _p(ws), nbytes, _stream()))
            return dx, dbeta, dgamma, None, None, None, None, None
        x = x.contiguous()  # no-op unless the forward ran channels-last natively and the backward cannot
        g = _f32c(g)
        dx = torch.empty_like(x)
        _lib.check(L.mmnc_gdn_backward_raw(_p(x), _p(g), B, C, HW, _p(beta), _p(gamma), bb, gb, ped, inverse, precision,
                                           _p(dx), _p(dbeta), _p(dgamma), _p(ws), nbytes, _stream()))
        if cl:
            dx = dx.contiguous(memory_format=torch.channels_last)
        return dx, dbeta, dgamma, None, None, None, None, None


def gdn_raw(x: Tensor, beta: Tensor, gamma: Tensor, beta_bound: float, gamma_bound: float, pedestal: float,
            inverse: bool, precision: str = "auto") -> Tensor:
    return _GDNRawFn.apply(x, beta, gamma, float(beta_bound), float(gamma_bound), float(pedestal), bool(inverse),
                           GDN_PRECISION[precision])


def gdn(x: Tensor, beta_eff: Tensor, gamma_eff: Tensor, inverse: bool, precision: str = "auto") -> Tensor:
    return _GDNFn.apply(x, beta_eff, gamma_eff, bool(inverse), GDN_PRECISION[precision])


# ------------------------------------------------------------------------------------------------ conv bias (f1)
def channel_sum(g: Tensor) -> Tensor:
    """(B, C, ...) NCHW -> (C,) sums over batch and space: the bias gradient of a convolution."""
    _need_cuda(g)
    g = _f32c(g)
    B, C, S = _bcs(g)
    out = torch.empty(C, dtype=torch.float32, device=g.device)
    ws = torch.empty(int(_lib.lib().mmnc_channel_sum_workspace_floats(B, C, S)), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().mmnc_channel_sum(_p(g), B, C, S, _p(ws), _p(out), _stream()))
    return out


class _BiasAddFn(torch.autograd.Function):
    """out + bias[None, :, None, None], in place on the convolution's output.  Forward is the element-wise add torch
    issues after a cuDNN convolution anyway; the backward replaces torch's generic reduction for the bias gradient
    with `mmnc_channel_sum` and passes the output gradient through untouched."""

    @staticmethod
    def forward(ctx, out, bias):
        ctx.mark_dirty(out)
        if out.is_contiguous() and out.dtype == torch.float32 and bias.dtype == torch.float32 and bias.is_contiguous():
            B, C, S = _bcs(out)
            _lib.check(_lib.lib().mmnc_bias_add(_p(out), _p(bias), B, C, S, _stream()))
        else:
            out.add_(bias.view(1, -1, *([1] * (out.dim() - 2))))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, (channel_sum(g) if ctx.needs_input_grad[1] else None)


def bias_add_(out: Tensor, bias: Tensor) -> Tensor:
    return _BiasAddFn.apply(out, bias)


# ------------------------------------------------------------------------------------------------ coding (a9-a12)
def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = 16) -> List[int]:
    """compressai._CXX.pmf_to_quantized_cdf look-alike (host function of the C ABI)."""
    n = len(pmf)
    arr = (ctypes.c_float * n)(*[float(v) for v in pmf])
    out = (ctypes.c_uint32 * (n + 1))()
    rc = _lib.lib().mmnc_pmf_to_quantized_cdf_h(arr, n, int(precision), out)
    if rc != 0:
        raise ValueError(_lib.lib().mmnc_last_error().decode())
    return list(out)


def build_indexes(scales: Tensor, scale_table: Tensor, scale_bound: float) -> Tensor:
    _need_cuda(scales, scale_table)
    scales = _f32c(scales.detach())
    table = _f32c(scale_table)
    out = torch.empty(scales.shape, dtype=torch.int32, device=scales.device)
    _lib.check(_lib.lib().mmnc_build_indexes(_p(scales), scales.numel(), _p(table), table.numel(), float(scale_bound),
                                             _p(out), _stream()))
    return out


class RansTables:
    """Device-side coding tables of one entropy model: CompressAI's `_quantized_cdf` / `_cdf_length` / `_offset`
    re-packed once (per update()) into the ragged uint16 table the kernels stage in shared memory."""

    def __init__(self, cdf: Tensor, cdf_sizes: Tensor, offsets: Tensor):
        _need_cuda(cdf, cdf_sizes, offsets)
        if cdf.dim() != 2 or cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        self.device = cdf.device
        cdf = cdf.int().contiguous()
        self.sizes = cdf_sizes.int().contiguous()
        self.offsets = offsets.int().contiguous()
        self.n_cdfs = int(cdf.shape[0])
        self.ragged_len = int(self.sizes.clamp(0, cdf.shape[1]).sum().item())  # once per update(), not per call
        cap = max(8, (self.ragged_len + 7) // 8 * 8)
        self.ragged = torch.zeros(cap, dtype=torch.int16, device=self.device)
        self.row_start = torch.empty(self.n_cdfs + 1, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().mmnc_rans_pack_tables(_p(cdf), _p(self.sizes), self.n_cdfs, int(cdf.shape[1]),
                                                    _p(self.row_start), _p(self.ragged), cap, _stream()))
        self.key = (cdf.data_ptr(), cdf._version, self.sizes.data_ptr(), self.offsets.data_ptr())


class _RansWorkspace:
    """Grow-only device scratch + pinned host staging, one per device: `compress` / `decompress` allocate nothing in
    steady state and move their results with one metadata copy and one payload copy."""

    _per_device = {}

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.upload_done = None  # CUDA event: the last asynchronous copy out of the pinned upload buffer
        self.bytes_per_stream = {}  # symbols per stream -> running estimate that sizes rans_encode's single D2H copy

    @classmethod
    def get(cls, device) -> "_RansWorkspace":
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        ws = cls._per_device.get(key)
        if ws is None:
            ws = cls._per_device[key] = cls(device)
        return ws

    def dev(self, name: str, nbytes: int) -> Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = self.bufs[name] = torch.empty(max(256, int(nbytes * 1.25)), dtype=torch.uint8, device=self.device)
        return t

    def host(self, name: str, nbytes: int) -> Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = self.bufs[name] = torch.empty(max(4096, int(nbytes * 1.25)), dtype=torch.uint8, pin_memory=True)
        return t


def rans_encode_device(symbols: Tensor, indexes: Optional[Tensor], channel_period: int, tables: RansTables):
    """The device part of `rans_encode`: -> (out, meta_bytes_aligned, n_streams) still on the GPU, nothing
    synchronised.  out = [meta | packed]: meta = int64 offsets[n + 1] | int32 nbytes[n] (padded to 16 bytes), packed =
    the strings back to back."""
    _need_cuda(symbols, indexes)
    n_streams = symbols.shape[0]
    sym = symbols.reshape(n_streams, -1)
    if sym.dtype != torch.int32 or not sym.is_contiguous():
        sym = sym.int().contiguous()
    n_sym = sym.shape[1]
    if indexes is not None:
        indexes = indexes.reshape(n_streams, -1)
        if indexes.dtype != torch.int32 or not indexes.is_contiguous():
            indexes = indexes.int().contiguous()
    L = _lib.lib()
    ws = _RansWorkspace.get(sym.device)
    slab_words = int(L.mmnc_rans_slab_words(n_sym))
    n1 = max(1, n_streams)
    staging = ws.dev("staging", n1 * max(1, n_sym) * 12)
    slabs = ws.dev("slabs", n1 * slab_words * 4)
    nbytes = ws.dev("nbytes", n1 * 4)
    meta_al = ((n1 + 1) * 8 + n1 * 4 + 15) // 16 * 16
    cap = n1 * slab_words * 4
    out = ws.dev("out", meta_al + cap)
    _lib.check(L.mmnc_rans_encode_batch(_p(sym), _p(indexes), int(channel_period), n_streams, n_sym, _p(tables.ragged),
                                        tables.ragged_len, _p(tables.row_start), _p(tables.sizes), _p(tables.offsets),
                                        tables.n_cdfs, _p(staging), _p(slabs), slab_words, _p(nbytes), _stream()))
    _lib.check(L.mmnc_rans_compact(_p(slabs), slab_words, _p(nbytes), n_streams, _p(out), _p(out[meta_al:]), cap,
                                   _stream()))
    return out, meta_al, n_streams


def rans_encode(symbols: Tensor, indexes: Optional[Tensor], channel_period: int, tables: RansTables) -> List[bytes]:
    """symbols (n_streams, ...) int32 on the GPU -> one CompressAI-format byte string per stream."""
    out, meta_al, n = rans_encode_device(symbols, indexes, channel_period, tables)
    if n == 0:
        return []
    ws = _RansWorkspace.get(symbols.device)
    stream = torch.cuda.current_stream(symbols.device)
    # ONE device-to-host copy in the common case: the metadata plus as many payload bytes as the previous calls needed
    # per stream (+ 25 %); a second copy fetches the remainder only when a batch codes unusually long strings
    n_sym = symbols[0].numel()
    guess = min(out.numel() - meta_al, int(n * ws.bytes_per_stream.get(n_sym, 0.6 * n_sym + 16) * 1.25) + 1024)
    host = ws.host("host_out", meta_al + guess)
    host[: meta_al + guess].copy_(out[: meta_al + guess], non_blocking=True)
    stream.synchronize()
    hv = host.numpy()
    offs_h = hv[: (n + 1) * 8].view("int64")
    nb_h = hv[(n + 1) * 8: (n + 1) * 8 + n * 4].view("int32")
    if (nb_h < 0).any():
        bad = [int(i) for i in (nb_h < 0).nonzero()[0][:8]]
        raise ValueError(f"rans_encode: malformed input for streams {bad} (index out of range or zero-width CDF bin)")
    total = int(offs_h[n])
    ws.bytes_per_stream[n_sym] = max(32.0, total / n)
    if total > guess:
        host2 = ws.host("host_out2", total)
        host2[:total].copy_(out[meta_al: meta_al + total], non_blocking=True)
        stream.synchronize()
        blob = host2.numpy()[:total].tobytes()
    else:
        blob = hv[meta_al: meta_al + total].tobytes()
    cuts = offs_h.tolist()
    return [blob[cuts[i]: cuts[i + 1]] for i in range(n)]


def rans_decode_device(blob_dev: Tensor, offs_dev: Tensor, lens_dev: Tensor, n_streams: int,
                       indexes: Optional[Tensor], channel_period: int, n_sym: int, tables: RansTables):
    """The device part of `rans_decode`: -> (symbols (n_streams, n_sym) int32, status (n_streams,) int32)."""
    dev = blob_dev.device
    if indexes is not None:
        indexes = indexes.reshape(n_streams, -1)
        if indexes.dtype != torch.int32 or not indexes.is_contiguous():
            indexes = indexes.int().contiguous()
        if indexes.shape[1] != n_sym:
            raise ValueError("indexes do not match the number of symbols")
    out = torch.empty((n_streams, n_sym), dtype=torch.int32, device=dev)
    status = torch.zeros(max(1, n_streams), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().mmnc_rans_decode_batch(_p(blob_dev), _p(offs_dev), _p(lens_dev), _p(indexes),
                                                 int(channel_period), n_streams, n_sym, _p(tables.ragged),
                                                 tables.ragged_len, _p(tables.row_start), _p(tables.sizes),
                                                 _p(tables.offsets), tables.n_cdfs, _p(out), _p(status), _stream()))
    return out, status


def rans_upload(strings: Sequence[bytes], device):
    """byte strings -> (blob, offsets int64, lengths int32) on the device with ONE pinned host-to-device copy.  Every
    stream starts at a multiple of 4 bytes (the kernel reads whole words)."""
    import numpy as np

    n = len(strings)
    lens = np.fromiter((len(s) for s in strings), dtype=np.int64, count=n)
    padded = (lens + 3) & ~3
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(padded, out=offs[1:])
    total = int(offs[n])
    head = (n + 1) * 8 + ((n * 4 + 7) // 8) * 8
    ws = _RansWorkspace.get(device)
    if ws.upload_done is not None:
        ws.upload_done.synchronize()  # the previous upload may still be reading the pinned buffer
    host = ws.host("host_up", head + max(total, 4))
    hv = host.numpy()
    hv[: (n + 1) * 8].view("int64")[:] = offs
    hv[(n + 1) * 8: (n + 1) * 8 + n * 4].view("int32")[:] = lens.astype(np.int32)
    if (padded == lens).all():
        hv[head: head + total] = np.frombuffer(b"".join(strings), dtype=np.uint8) if total else 0
    else:  # malformed lengths: place each stream at its aligned offset
        for i, s_ in enumerate(strings):
            hv[head + offs[i]: head + offs[i] + len(s_)] = np.frombuffer(s_, dtype=np.uint8)
    dev_buf = ws.dev("dev_up", head + max(total, 4))
    dev_buf[: head + max(total, 4)].copy_(host[: head + max(total, 4)], non_blocking=True)
    ws.upload_done = torch.cuda.Event()
    ws.upload_done.record(torch.cuda.current_stream(dev_buf.device))
    offs_dev = dev_buf[: (n + 1) * 8].view(torch.int64)
    lens_dev = dev_buf[(n + 1) * 8: (n + 1) * 8 + n * 4].view(torch.int32) if n else dev_buf[:0].view(torch.int32)
    return dev_buf[head:], offs_dev, lens_dev


def rans_decode(strings: Sequence[bytes], indexes: Optional[Tensor], channel_period: int, n_sym: int,
                tables: RansTables) -> Tensor:
    """byte strings -> (n_streams, n_sym) int32 symbols on the GPU."""
    _need_cuda(indexes)
    n_streams = len(strings)
    dev = tables.device
    blob_dev, offs_dev, lens_dev = rans_upload(strings, dev)
    out, status = rans_decode_device(blob_dev, offs_dev, lens_dev, n_streams, indexes, channel_period, n_sym, tables)
    if n_streams:
        st = status[:n_streams].cpu()  # also the point where the pinned upload buffer may be reused
        if bool((st != 0).any()):
            bad = [(int(i), int(st[i])) for i in (st != 0).nonzero().reshape(-1)[:8]]
            raise ValueError(f"rans_decode: corrupt or truncated stream(s) {bad}")
    return out
