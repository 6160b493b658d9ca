"""Drop-in `GDN` (CompressAI 1.2.4 `compressai.layers.GDN`) and the conv helpers the reference imports.

Reference sites: /root/reference/src/models/multi_task_compressor.py:18-19 (imports), 144-173 (GDN(C),
GDN(C, inverse=True) in the heads), /root/reference/src/models/disjoint_latent.py:147-158.
State-dict keys match CompressAI: `beta`, `gamma`, `beta_reparam.pedestal`, `beta_reparam.lower_bound.bound`,
`gamma_reparam.pedestal`, `gamma_reparam.lower_bound.bound`.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .entropy_models import LowerBound


class _Conv2d(nn.Conv2d):
    """nn.Conv2d (same parameters, same state-dict keys, the convolution itself on cuDNN) whose bias gradient is
    reduced by `mmnc_channel_sum` instead of torch's generic reduction: on CUDA the bias is added by a tiny autograd
    function after the bias-free convolution - which is what torch does internally after a cuDNN convolution."""

    def forward(self, x: Tensor) -> Tensor:
        if self.bias is None or not x.is_cuda or x.dtype != torch.float32 or ops._is_channels_last(x):
            return super().forward(x)
        return ops.bias_add_(self._conv_forward(x, self.weight, None), self.bias)


class _ConvTranspose2d(nn.ConvTranspose2d):
    def forward(self, x: Tensor, output_size=None) -> Tensor:
        if self.bias is None or not x.is_cuda or x.dtype != torch.float32 or ops._is_channels_last(x):
            return super().forward(x, output_size)
        output_padding = self._output_padding(x, output_size, self.stride, self.padding, self.kernel_size, 2,
                                              self.dilation)
        out = F.conv_transpose2d(x, self.weight, None, self.stride, self.padding, output_padding, self.groups,
                                 self.dilation)
        return ops.bias_add_(out, self.bias)


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """compressai.models.utils.conv — stays on cuDNN (out of the rate path)."""
    return _Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    """compressai.models.utils.deconv — stays on cuDNN."""
    return _ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                            output_padding=stride - 1, padding=kernel_size // 2)


class NonNegativeParametrizer(nn.Module):
    """max(p, bound)^2 - pedestal with LowerBound's custom gradient, as one kernel each way."""

    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)
        self._pedestal_f = float(torch.tensor(pedestal, dtype=torch.float32))

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        return ops.nonneg_reparam(x, self.lower_bound.value(), self._pedestal_f)


class GDN(nn.Module):
    r"""Generalized Divisive Normalization: y_i = x_i / sqrt(beta_i + sum_j gamma_ij x_j^2) (or its inverse).

    `precision` selects the channel contraction: "auto" (tensor cores where the shape allows, fp32 SIMT
    otherwise), "fp32", "tf32", "3xtf32".
    """

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1,
                 precision: str = "auto"):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)
        self.precision = precision
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 4:
            raise ValueError(f"GDN expects (B, C, H, W), got {tuple(x.shape)}")
        # NonNegativeParametrizer (both parameters share the pedestal 2^-36) is applied inside the kernels
        return ops.gdn_raw(x, self.beta, self.gamma, self.beta_reparam.lower_bound.value(),
                           self.gamma_reparam.lower_bound.value(), self.beta_reparam._pedestal_f, self.inverse,
                           self.precision)
