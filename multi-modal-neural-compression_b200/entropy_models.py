"""Drop-in `EntropyModel`, `EntropyBottleneck` and `GaussianConditional` (CompressAI 1.2.4 interfaces).

Same constructor arguments, method names, return values, error behaviour and state-dict keys as the CompressAI
modules the reference uses (reference call sites
/root/reference/src/models/multi_task_compressor.py:387, 487-488, 495, 509, 543-546) — but `forward`,
`compress`, `decompress`, `build_indexes` and `loss` run as hand-written sm_100a kernels behind the C ABI
(include/mmnc_b200.h).  CUDA only: a CPU tensor raises.

What stays in stock torch ops, on purpose: building the integer CDF tables in `update()`.  The reference calls it
once after loading weights, on the CPU, before `.to(device)` (/root/reference/src/compress.py:101-105); the
tables decide every bitstream, so the fp32 pmf is evaluated with the same torch CPU ops CompressAI uses and only
the integer part (`pmf_to_quantized_cdf`) is native (csrc/cdf_host.cpp).  The kernels consume tables, they never
produce them.
"""
from __future__ import annotations

from typing import List, Optional

import math
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops


class LowerBound(nn.Module):
    """Kept for state-dict compatibility (`...lower_bound.bound` buffers); the bound itself is applied in-kernel."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._bound_f = float(torch.tensor(float(bound), dtype=torch.float32))

    def value(self) -> float:
        return self._bound_f

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._bound_f = float(self.bound.detach().cpu().reshape(-1)[0])

    def forward(self, x: Tensor) -> Tensor:  # generic use outside the fused kernels
        return _LowerBoundFn.apply(x, self.bound)


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        if entropy_coder not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (available: ans)')
        self.entropy_coder_precision = int(entropy_coder_precision)
        if self.entropy_coder_precision != 16:
            raise NotImplementedError("the rANS kernels implement CompressAI's 16-bit precision only")
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # CompressAI resizes these registered buffers on load (update_registered_buffers); same here so that
        # state dicts saved after update() load into a freshly built module.
        self.__dict__.pop("_rans_tables_cache", None)
        for name in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table"):
            key = prefix + name
            buf = getattr(self, name, None)
            if key in state_dict and isinstance(buf, torch.Tensor) and buf.shape != state_dict[key].shape:
                setattr(self, name, torch.empty(state_dict[key].shape, dtype=buf.dtype, device=buf.device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _lik_bound(self) -> float:
        return self.likelihood_lower_bound.value() if self.use_likelihood_bound else 0.0

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    # ------------------------------------------------------------------ quantise (a2)
    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return ops._AddNoise.apply(inputs, None, None, 0)  # d/dx = 1, like x + noise
        if mode == "dequantize":
            return ops._StraightRound.apply(inputs, means)
        return ops.quantize_symbols(inputs.detach(), None if means is None else means.detach())

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            return ops.dequantize_symbols(inputs, means.detach()).type_as(means)
        return ops.dequantize_symbols(inputs, None).type(dtype)

    # ------------------------------------------------------------------ tables (a9)
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=pmf.device)
        pmf_h, tail_h, len_h = pmf.detach().cpu(), tail_mass.detach().cpu(), pmf_length.cpu().tolist()
        for i in range(len(len_h)):
            prob = torch.cat((pmf_h[i, : len_h[i]], tail_h[i]), dim=0)
            row = torch.IntTensor(ops.pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision))
            cdf[i, : row.size(0)] = row
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    # ------------------------------------------------------------------ coding (a11, a12)
    def _rans_tables(self) -> "ops.RansTables":
        """The ragged 16-bit coding tables the kernels keep in shared memory, rebuilt only when update(), .to() or a
        state-dict load replaced the CDF buffers."""
        cdf = self._quantized_cdf
        key = (cdf.data_ptr(), cdf._version, self._cdf_length.data_ptr(), self._offset.data_ptr())
        cached = self.__dict__.get("_rans_tables_cache")
        if cached is None or cached.key != key:
            cached = ops.RansTables(cdf, self._cdf_length, self._offset)
            if cached.key != key:  # the buffers were int32 and contiguous already, so no copy was made; keep honest
                cached.key = key
            self.__dict__["_rans_tables_cache"] = cached
        return cached

    def _drop_rans_tables(self):
        self.__dict__.pop("_rans_tables_cache", None)

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .cpu(): the buffers move, the packed tables do not
        self._drop_rans_tables()
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_rans_tables_cache", None)  # device scratch is not part of the module's state
        return state

    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None) -> List[bytes]:
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        symbols = self.quantize(inputs, "symbols", means)
        return ops.rans_encode(symbols, indexes, 0, self._rans_tables())

    def decompress(self, strings, indexes: Tensor, dtype: torch.dtype = torch.float,
                   means: Optional[Tensor] = None) -> Tensor:
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        n_sym = int(indexes[0].numel()) if indexes.size(0) > 0 else 0
        symbols = ops.rans_decode(strings, indexes, 0, n_sym, self._rans_tables())
        return self.dequantize(symbols.reshape(indexes.size()), means, dtype)


class EntropyBottleneck(EntropyModel):
    """Factorised-prior entropy bottleneck; `forward(x, training=None) -> (x_hat, likelihood)`."""

    _offset: Tensor

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters=(3, 3, 3, 3), likelihood_form: str = "sign", **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("the EB kernels are specialised for CompressAI's default filters=(3,3,3,3)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        if likelihood_form not in ops.EB_FORM:
            raise ValueError(f"likelihood_form must be one of {sorted(ops.EB_FORM)}")
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
        # filled by forward(): per-channel sum over (batch, space) of ln(likelihood) of the last call
        self.last_log_likelihood_sums: Optional[Tensor] = None

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def packed_parameters(self) -> Tensor:
        """(C, 58) raw parameters in the kernel's per-channel order (differentiable torch.cat)."""
        C, parts = self.channels, []
        for i in range(len(self.filters) + 1):
            parts.append(getattr(self, f"_matrix{i:d}").reshape(C, -1))
            parts.append(getattr(self, f"_bias{i:d}").reshape(C, -1))
            if i < len(self.filters):
                parts.append(getattr(self, f"_factor{i:d}").reshape(C, -1))
        return torch.cat(parts, dim=1)

    # ------------------------------------------------------------------ forward (a3)
    def forward(self, x: Tensor, training: Optional[bool] = None, noise: Optional[Tensor] = None):
        if training is None:
            training = self.training
        if x.dim() < 2 or x.size(1) != self.channels:
            raise ValueError(f"expected (B, {self.channels}, ...), got {tuple(x.shape)}")
        outputs, likelihood, lnsum = ops.entropy_bottleneck_forward(
            x, self.packed_parameters(), self._get_medians().reshape(-1), training, self._lik_bound(),
            self.likelihood_form, noise=noise)
        self.last_log_likelihood_sums = lnsum
        return outputs, likelihood

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        if not stop_gradient:
            raise NotImplementedError("differentiable _logits_cumulative is fused into forward()")
        if inputs.is_cuda:
            return ops.eb_logits(inputs, self.packed_parameters())
        return self._logits_cumulative_torch(inputs)

    def _logits_cumulative_torch(self, inputs: Tensor) -> Tensor:
        """Stock torch ops, used only by update() to build tables (see module docstring)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            logits = torch.matmul(F.softplus(getattr(self, f"_matrix{i:d}").detach()), logits)
            logits = logits + getattr(self, f"_bias{i:d}").detach()
            if i < len(self.filters):
                logits = logits + torch.tanh(getattr(self, f"_factor{i:d}").detach()) * torch.tanh(logits)
        return logits

    # ------------------------------------------------------------------ aux loss (a4)
    def loss(self) -> Tensor:
        return ops.eb_aux_loss(self.quantiles, self.packed_parameters(), self.target)

    # ------------------------------------------------------------------ tables (a9)
    @torch.no_grad()
    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length, device=pmf_start.device)
        samples = samples[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative_torch(samples - 0.5)
        upper = self._logits_cumulative_torch(samples + 0.5)
        if self.likelihood_form == "sign":
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        else:
            pmf = torch.sigmoid(upper) - torch.sigmoid(lower)
        pmf = pmf[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._cdf_length = pmf_length + 2
        self._drop_rans_tables()
        return True

    # ------------------------------------------------------------------ coding (a11, a12)
    def _build_indexes(self, size):
        dims = len(size)
        N, C = size[0], size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(C, device=self._quantized_cdf.device).view(*view_dims)
        return indexes.int().repeat(N, 1, *size[2:])

    def compress(self, x: Tensor) -> List[bytes]:
        if x.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if x.size(1) != self._quantized_cdf.size(0):
            raise ValueError("`inputs` and `indexes` should have the same size.")
        # indexes are the channel ids: implicit in the kernel (position // spatial_size), never materialised
        symbols = ops.quantize_symbols(x.detach(), self._get_medians().detach().reshape(1, -1, *([1] * (x.dim() - 2))))
        period = int(x[0, 0].numel())
        return ops.rans_encode(symbols, None, period, self._rans_tables())

    def decompress(self, strings, size) -> Tensor:
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        C = self._quantized_cdf.size(0)
        output_size = (len(strings), C, *size)
        period = 1
        for d in size:
            period *= int(d)
        symbols = ops.rans_decode(strings, None, period, C * period, self._rans_tables())
        medians = self._get_medians().detach().reshape(1, -1, *([1] * len(size)))
        return ops.dequantize_symbols(symbols.reshape(output_size), medians)


class GaussianConditional(EntropyModel):
    """Zero-mean (or given-mean) Gaussian conditional; `forward(inputs, scales, means=None, training=None)`."""

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
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self.last_log_likelihood_sums: Optional[Tensor] = None

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs: Tensor) -> Tensor:
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        # scipy.stats.norm.ppf without scipy: inverse error function in float64
        return float(math.sqrt(2.0) * torch.erfinv(torch.tensor(2.0 * quantile - 1.0, dtype=torch.float64)))

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    @torch.no_grad()
    def update(self):
        """Stock torch ops on the module's device (CPU in the reference flow), see module docstring."""
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        device = pmf_center.device
        samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None])
        samples_scale = self.scale_table.unsqueeze(1).float()
        samples = samples.float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2
        self._drop_rans_tables()

    # ------------------------------------------------------------------ forward (a5)
    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None):
        if training is None:
            training = self.training
        outputs, likelihood, lnsum = ops.gaussian_conditional_forward(
            inputs, scales, means, training, self.lower_bound_scale.value(), self._lik_bound(), noise=noise)
        self.last_log_likelihood_sums = lnsum
        return outputs, likelihood

    # ------------------------------------------------------------------ indexes (a10)
    def build_indexes(self, scales: Tensor) -> Tensor:
        if self.scale_table.numel() == 0:
            raise ValueError("Uninitialized scale table. Run update_scale_table() first")
        return ops.build_indexes(scales, self.scale_table, self.lower_bound_scale.value())
