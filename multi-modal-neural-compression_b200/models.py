"""Drop-in `ScaleHyperprior(N, M)` (CompressAI 1.2.4 `compressai.models.ScaleHyperprior`).

What the reference relies on: `backbone_class(N=, M=)`, item assignment into `g_a` / `g_s`
(/root/reference/src/models/multi_task_compressor.py:186-191), attribute replacement `.g_s = DummyModule()`
(disjoint_latent.py:179), `model(x)["x_hat"]`, `["likelihoods"]["y"|"z"]` (mtc.py:495-500),
`.compress(x)["strings"]` (mtc.py:509-516) and the sub-modules `entropy_bottleneck`, `gaussian_conditional`,
`h_s`, `g_s` (mtc.py:543-547).  Convolutions stay on torch/cuDNN; everything else is this package's kernels.

Extra (non-breaking) output: `forward` also returns `"log_likelihood_sums": {"y": (M,), "z": (N,)}` — the
per-channel sums of ln(likelihood) already reduced inside the likelihood kernels, so that the RD loss does not
have to read the likelihood tensors again.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import GDN, conv, deconv

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):  # noqa: A002
    """compressai.models.base.get_scale_table (/root/reference/src/models/multi_task_compressor.py:20, 487)."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class ScaleHyperprior(nn.Module):
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

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        x_hat = self.g_s(y_hat)
        return {
            "x_hat": x_hat,
            "likelihoods": {"y": y_likelihoods, "z": z_likelihoods},
            "log_likelihood_sums": {"y": self.gaussian_conditional.last_log_likelihood_sums,
                                    "z": self.entropy_bottleneck.last_log_likelihood_sums},
        }

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
        # CompressAI decodes the strings it has just produced to obtain z_hat; the decoder returns round(z - median) +
        # median symbol by symbol, i.e. exactly the eval-mode quantisation of z (asserted bit for bit in
        # tests/test_gpu_parity.py::test_eb_compress_bit_exact), so the decode round trip (H2D of the strings, a
        # decode pass and a host synchronisation) is skipped.
        z_hat = self.entropy_bottleneck.quantize(
            z, "dequantize", self.entropy_bottleneck._get_medians().detach().reshape(1, -1, *([1] * (z.dim() - 2))))
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
