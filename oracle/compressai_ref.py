"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  Stock-torch restatement (CPU or any device, fp32 or fp64) of the CompressAI 1.2.4 modules
the reference's rate path runs through.  CompressAI itself (`requirements.txt:10`) is not vendored, not
installed and not fetchable here, so this follows the published algorithm in
CompressAI's own op order, anchored on the reference call sites:

  GDN / conv / deconv / get_scale_table   /root/reference/src/models/multi_task_compressor.py:18-20,144-173
  ScaleHyperprior(N=, M=), g_a[0]/g_s[-1]  /root/reference/src/models/multi_task_compressor.py:186-191
  model(x)["x_hat"|"likelihoods"]          /root/reference/src/models/multi_task_compressor.py:495-500
  .compress(x)["strings"]                  /root/reference/src/models/multi_task_compressor.py:509-516
  EB.decompress / h_s / build_indexes /
  GC.decompress / g_s                      /root/reference/src/models/multi_task_compressor.py:543-547
  entropy_bottleneck.loss()                /root/reference/src/models/multi_task_compressor.py:386-387
  update_scale_table / update              /root/reference/src/models/multi_task_compressor.py:486-489

Parameter and buffer names follow CompressAI so state dicts interchange with the product
modules in `multi-modal-neural-compression_b200/`.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import native


# ----------------------------------------------------------------------------------------------- A.0
def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def get_scale_table(min=0.11, max=256, levels=64):  # noqa: A002 - CompressAI's argument names
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


# ----------------------------------------------------------------------------------------------- A.2
class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        passthrough = (x >= bound) | (grad_output < 0)
        return passthrough * grad_output, None


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound.to(x.dtype))


# ----------------------------------------------------------------------------------------------- A.5
class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal.to(x.dtype)


class GDN(nn.Module):
    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        C = x.size(1)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


# ----------------------------------------------------------------------------------------------- A.1 / A.6
class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder=None, entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._enc = native.RansEncoder()
        self._dec = native.RansDecoder()
        # "faithful" = five .tolist() marshals per image like CompressAI; "lean" = ndarray views
        self.marshalling = "faithful"

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # CompressAI resizes these registered buffers on load (update_registered_buffers); same here so that
        # state dicts saved after update() load into a freshly built module.
        for name in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table"):
            key = prefix + name
            buf = getattr(self, name, None)
            if key in state_dict and isinstance(buf, torch.Tensor) and buf.shape != state_dict[key].shape:
                setattr(self, name, torch.empty(state_dict[key].shape, dtype=buf.dtype, device=buf.device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
            return inputs + noise
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=pmf.device)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = torch.IntTensor(native.pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision))
            cdf[i, : row.size(0)] = row
        return cdf

    def _check_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if self._cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None) -> List[bytes]:
        symbols = self.quantize(inputs, "symbols", means)
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_tables()
        strings = []
        if self.marshalling == "faithful":
            for i in range(symbols.size(0)):
                strings.append(self._enc.encode_with_indexes(
                    symbols[i].reshape(-1).int().tolist(), indexes[i].reshape(-1).int().tolist(),
                    self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                    self._offset.reshape(-1).int().tolist()))
        else:
            cdf, ln, off = (self._quantized_cdf.cpu().numpy(), self._cdf_length.cpu().numpy(),
                            self._offset.cpu().numpy())
            s_np, i_np = symbols.cpu().numpy(), indexes.int().cpu().numpy()
            for i in range(symbols.size(0)):
                strings.append(native.encode_with_indexes_np(s_np[i], i_np[i], cdf, ln, off))
        return strings

    def decompress(self, strings, indexes: Tensor, dtype: torch.dtype = torch.float,
                   means: Optional[Tensor] = None) -> Tensor:
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        if means is not None and means.size()[:2] != indexes.size()[:2]:
            raise ValueError("Invalid means or indexes parameters")
        cdf = self._quantized_cdf
        outputs = cdf.new_empty(indexes.size())
        if self.marshalling == "faithful":
            for i, s in enumerate(strings):
                values = self._dec.decode_with_indexes(
                    s, indexes[i].reshape(-1).int().tolist(), cdf.tolist(),
                    self._cdf_length.reshape(-1).int().tolist(), self._offset.reshape(-1).int().tolist())
                outputs[i] = torch.tensor(values, device=outputs.device, dtype=outputs.dtype).reshape(
                    outputs[i].size())
        else:
            cdf_np, ln, off = cdf.cpu().numpy(), self._cdf_length.cpu().numpy(), self._offset.cpu().numpy()
            i_np = indexes.int().cpu().numpy()
            for i, s in enumerate(strings):
                values = native.decode_with_indexes_np(s, i_np[i], cdf_np, ln, off)
                outputs[i] = torch.from_numpy(values).reshape(outputs[i].size())
        return self.dequantize(outputs, means, dtype)


# ----------------------------------------------------------------------------------------------- A.3
class EntropyBottleneck(EntropyModel):
    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters=(3, 3, 3, 3), likelihood_form: str = "sign", **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        # "sign": CompressAI 1.2.x `sign = -sign(lower+upper)` form; "plain": sigmoid(upper)-sigmoid(lower)
        self.likelihood_form = likelihood_form

        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))

        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood_from_logits(self, lower, upper):
        if self.likelihood_form == "sign":
            sign = -torch.sign(lower + upper)
            sign = sign.detach()
            return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        return torch.sigmoid(upper) - torch.sigmoid(lower)

    def _likelihood(self, inputs: Tensor) -> Tensor:
        lower = self._logits_cumulative(inputs - 0.5, stop_gradient=False)
        upper = self._logits_cumulative(inputs + 0.5, stop_gradient=False)
        return self._likelihood_from_logits(lower, upper)

    def forward(self, x: Tensor, training: Optional[bool] = None, noise: Optional[Tensor] = None):
        """`noise` (test hook, same shape as x) replaces the uniform_ draw when training."""
        if training is None:
            training = self.training
        perm = list(range(x.dim()))
        perm[0], perm[1] = 1, 0
        inv_perm = list(np.argsort(perm))
        x = x.permute(*perm).contiguous()
        shape = x.size()
        values = x.reshape(x.size(0), 1, -1)
        if training and noise is not None:
            outputs = values + noise.permute(*perm).contiguous().reshape(values.shape)
        else:
            outputs = self.quantize(values, "noise" if training else "dequantize", self._get_medians())
        likelihood = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*inv_perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*inv_perm).contiguous()
        return outputs, likelihood

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = pmf_length.max().item()
        samples = torch.arange(max_length, device=pmf_start.device)
        samples = samples[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
        pmf = self._likelihood_from_logits(lower, upper)[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        with torch.no_grad():
            self._quantized_cdf = self._pmf_to_cdf(pmf.detach(), tail_mass.detach(), pmf_length, max_length)
        self._cdf_length = pmf_length + 2
        return True

    def _build_indexes(self, size):
        dims = len(size)
        N, C = size[0], size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(C).view(*view_dims)
        return indexes.int().repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        indexes = self._build_indexes(x.size())
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians.dtype, medians)


# ----------------------------------------------------------------------------------------------- A.4
class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        device = pmf_center.device
        samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None])
        samples_scale = self.scale_table.unsqueeze(1)
        samples = samples.float()
        samples_scale = samples_scale.float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None):
        if training is None:
            training = self.training
        if training and noise is not None:
            outputs = inputs + noise
        else:
            outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales: Tensor) -> Tensor:
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes


# ----------------------------------------------------------------------------------------------- A.0
class ScaleHyperprior(nn.Module):
    """`compressai.models.ScaleHyperprior(N, M)` with pluggable layer classes (used by the oracle wrappers)."""

    def __init__(self, N, M, **kwargs):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= self.entropy_bottleneck.update(force=force)
        return updated

    def compress(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_strings = self.gaussian_conditional.compress(y, indexes)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}
