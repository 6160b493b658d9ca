"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED (CompressAI absent).

Plain-torch restatement (no Lightning, no W&B) of the reference's four multi-task compressors: topology,
forward, the rate-distortion loss formulas, compress/decompress and the two-optimizer training step.  Built on
oracle/compressai_ref.py.  Each piece cites the reference lines it follows.

  topology of the heads        /root/reference/src/models/multi_task_compressor.py:109-177
  backbone surgery             /root/reference/src/models/multi_task_compressor.py:179-193
  Mixed (-m 2) / Single (-m 1) /root/reference/src/models/mixed_latent.py:70-162, single_task_compressor.py:13-55
  Disjoint (-m 3)              /root/reference/src/models/disjoint_latent.py:40-194
  Shared (-m 4)                /root/reference/src/models/shared_latent.py:21-162
  distortion terms             /root/reference/src/models/multi_task_compressor.py:223-276
  bits per pixel               /root/reference/src/models/multi_task_compressor.py:278-293, 302-357
  uncertainty weighting        /root/reference/src/loss_balancing.py:21-54
  train step                   /root/reference/src/models/multi_task_compressor.py:399-476
  compress / decompress        /root/reference/src/models/multi_task_compressor.py:507-549
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import compressai_ref as cai

# /root/reference/src/datasets/task_configs.py:7-33 (the constants the loss needs)
TASKS = {
    "rgb": dict(in_channels=3, out_channels=3, loss="mse"),
    "depth_euclidean": dict(in_channels=1, out_channels=1, loss="mse"),
    "normal": dict(in_channels=3, out_channels=3, loss="mse"),
    "semantic": dict(in_channels=1, out_channels=17, loss="cross-entropy"),
    "mono": dict(in_channels=1, out_channels=1, loss="mse"),
}


def encoder_head(cin: int, cout: int, L=cai) -> nn.Sequential:
    """mtc.py:160-173 — conv3 + GDN then five stride-2 conv5 + GDN."""
    mid = cout // 2
    layers: List[nn.Module] = [L.conv(cin, mid, kernel_size=3, stride=1), L.GDN(mid), L.conv(mid, cout), L.GDN(cout)]
    for _ in range(4):
        layers += [L.conv(cout, cout), L.GDN(cout)]
    return nn.Sequential(*layers)


def decoder_head(cin: int, cout: int, L=cai) -> nn.Sequential:
    """mtc.py:144-158 — four deconvs and three conv3 with IGDN in between."""
    mid = cin // 2
    ig = lambda c: L.GDN(c, inverse=True)  # noqa: E731
    return nn.Sequential(
        L.deconv(cin, mid), ig(mid), L.conv(mid, mid, kernel_size=3, stride=1), ig(mid),
        L.deconv(mid, mid), ig(mid), L.conv(mid, mid, kernel_size=3, stride=1), ig(mid),
        L.deconv(mid, cout), ig(cout), L.deconv(cout, cout), ig(cout),
        L.conv(cout, cout, kernel_size=3, stride=1))


def upsampled_decoder_head(cin: int, conv_channels: int, n_tasks: int, cout: int, L=cai) -> nn.Sequential:
    """disjoint.py:139-160 — four extra deconvs replacing the removed g_s, then the ordinary decoder head."""
    w = conv_channels // n_tasks
    ig = lambda c: L.GDN(c, inverse=True)  # noqa: E731
    return nn.Sequential(L.deconv(cin, w), ig(w), L.deconv(w, w), ig(w), L.deconv(w, w), ig(w),
                         L.deconv(w, conv_channels), decoder_head(conv_channels, cout, L))


class Identity(nn.Module):  # utils.py:56-61 DummyModule
    def forward(self, x):
        return x


class UncertaintyWeights(nn.Module):
    """loss_balancing.py:21-54: exp(-s_t) * L_t + s_t, zeroed where L_t == 0."""

    def __init__(self, n: int):
        super().__init__()
        self.log_vars = nn.Parameter(torch.zeros(n))

    def forward(self, losses: torch.Tensor) -> torch.Tensor:
        mask = losses != 0.0
        return (torch.exp(-self.log_vars) * losses + self.log_vars) * mask


class ReferenceCompressor(nn.Module):
    """kind: 1 single, 2 mixed, 3 disjoint, 4 shared (the reference's `-m` flag, src/train.py:89-99)."""

    def __init__(self, kind: int, tasks: Sequence[str], latent_channels: int, conv_channels: int,
                 lmbda: float = 1.0, learning_rate_main: float = 1e-5, learning_rate_aux: float = 1e-3, layers=cai):
        super().__init__()
        self.kind, self.tasks, self.T = int(kind), tuple(tasks), len(tasks)
        self.lmbda, self.lr_main, self.lr_aux = float(lmbda), learning_rate_main, learning_rate_aux
        L, T, c = layers, self.T, int(conv_channels)
        cin = [TASKS[t]["in_channels"] for t in tasks]
        cout = [TASKS[t]["out_channels"] for t in tasks]
        if kind == 1:
            assert T == 1
        if kind == 4 and latent_channels % (T + 1):          # shared.py:34-44
            latent_channels = latent_channels // (T + 1) * (T + 1)
        self.M = int(latent_channels)                         # what the backbone is built with (B4)
        self.group = {1: self.M, 2: self.M, 3: self.M // T, 4: self.M // (T + 1)}[kind]
        N = c * T
        model = nn.ModuleDict()
        model["input_heads"] = nn.ModuleList([encoder_head(cin[i], c, L) for i in range(T)])
        backbone = L.ScaleHyperprior(N=N, M=self.M)
        backbone.g_a[0] = L.conv(N, N)                        # mtc.py:190-191
        backbone.g_s[-1] = L.deconv(N, N)
        if kind in (3, 4):
            backbone.g_s = Identity()                         # disjoint.py:179, shared.py:73
        model["compressor"] = backbone
        if kind in (1, 2):
            heads = [decoder_head(N, cout[i], L) for i in range(T)]
        else:
            width = self.group if kind == 3 else 2 * self.group
            heads = [upsampled_decoder_head(width, c, T, cout[i], L) for i in range(T)]
        model["output_heads"] = nn.ModuleList(heads)
        self.model = model
        self.loss_balancer = UncertaintyWeights(T) if kind != 1 else None   # single.py:55 (identity)

    # ------------------------------------------------------------------ forward
    def forward_input_heads(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        return torch.cat([self.model["input_heads"][i](batch[t]) for i, t in enumerate(self.tasks)], dim=1)

    def _task_slice(self, t: torch.Tensor, i: int) -> torch.Tensor:
        return t[:, i * self.group:(i + 1) * self.group]

    def forward_output_heads(self, lat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for i, task in enumerate(self.tasks):
            if self.kind in (1, 2):
                inp = lat                                                     # mixed.py:155-162
            elif self.kind == 3:
                inp = self._task_slice(lat, i)                                # disjoint.py:187-194
            else:                                                             # shared.py:149-162
                B, _, H, W = lat.shape
                inp = torch.stack([self._task_slice(lat, i), lat[:, -self.group:]], dim=1).reshape(B, -1, H, W)
            out[task] = self.model["output_heads"][i](inp)
        return out

    def forward(self, batch):
        o = self.model["compressor"](self.forward_input_heads(batch))
        return self.forward_output_heads(o["x_hat"]), o["likelihoods"]

    # ------------------------------------------------------------------ losses
    @staticmethod
    def reconstruction_loss(x_hat, x, kind: str):
        if kind == "mse":
            return F.mse_loss(x, x_hat, reduction="none").sum(dim=[1, 2, 3]).mean(dim=[0]) / x.shape[1]
        if kind == "l1":
            return F.l1_loss(x, x_hat, reduction="none").sum(dim=[1, 2, 3]).mean(dim=[0]) / x.shape[1]
        if kind == "cross-entropy":
            return F.cross_entropy(input=x_hat, target=x.squeeze(1).long(), reduction="mean")
        raise NotImplementedError(kind)

    def multitask_reconstruction_loss(self, x, x_hats, log_dir="train"):
        logs, raw = {}, []
        for task in self.tasks:
            name = TASKS[task]["loss"]
            raw.append(self.reconstruction_loss(x_hats[task], x[task], name))
            logs[f"{log_dir}/{task}/{name}"] = raw[-1]
        raw_t = torch.stack(raw)
        if self.loss_balancer is not None:
            weighted = self.loss_balancer(raw_t)
            for i, task in enumerate(self.tasks):
                logs[f"uncertainty-weight/{task}"] = self.loss_balancer.log_vars[i]
        else:
            weighted = raw_t
        return weighted.sum(), logs

    @staticmethod
    def bits_per_pixel(lik: torch.Tensor, num_pixels: int):
        v = torch.log(lik).sum()
        v = v / -torch.log(torch.tensor(2.0))
        return v / num_pixels

    def multitask_compression_loss(self, lik, x_hats, log_dir="train"):
        logs = {}
        px = lambda task: x_hats[task].shape[0] * x_hats[task].shape[2] * x_hats[task].shape[3]  # noqa: E731
        z_bpp = self.bits_per_pixel(lik["z"], px(self.tasks[0]))
        if self.kind in (1, 2):                                               # mixed.py:70-118
            y_bpp = self.bits_per_pixel(lik["y"], px(self.tasks[0]))
            for task in self.tasks:
                logs[f"{log_dir}/{task}/compression_loss"] = y_bpp + z_bpp
            return (y_bpp + z_bpp) / self.T, logs
        total = 0.0                                                           # mtc.py:302-357
        for i, task in enumerate(self.tasks):
            b = self.bits_per_pixel(self._task_slice(lik["y"], i), px(task))
            logs[f"{log_dir}/{task}/compression_loss"] = b + z_bpp
            total = total + b
        total = (total + z_bpp) / self.T
        if self.kind == 4:                                                    # shared.py:118-147
            s = self.bits_per_pixel(lik["y"][:, -self.group:], px(self.tasks[0]))
            logs[f"{log_dir}/shared/compression_loss"] = s + z_bpp
            total = total + s / self.T
        return total, logs

    def auxiliary_loss(self):
        return self.model["compressor"].entropy_bottleneck.loss()

    def average_metrics(self, x, x_hats, log_dir):                           # mtc.py:359-384
        from . import metrics_ref
        return metrics_ref.average_metrics(self.tasks, x, x_hats, log_dir)

    def rd_loss(self, batch, log_dir="train", with_metrics=False):
        x_hats, lik = self.forward(batch)
        rec, l1 = self.multitask_reconstruction_loss(batch, x_hats, log_dir)
        comp, l2 = self.multitask_compression_loss(lik, x_hats, log_dir)
        loss = self.lmbda * rec + comp
        logs = {f"{log_dir}/rec_loss": rec, f"{log_dir}/compression_loss": comp, f"{log_dir}/loss": loss}
        logs.update(l1)
        logs.update(l2)
        if with_metrics:                                                      # mtc.py:468: every step, train and val
            self._last = (batch, x_hats)
            if log_dir != "train":
                logs.update(self.average_metrics(batch, x_hats, log_dir))
        return loss, logs

    # ------------------------------------------------------------------ optimisation (mtc.py:389-466)
    def configure_optimizers(self, total_steps: int = 1000):
        main = [p for n, p in self.model.named_parameters() if not n.endswith(".quantiles")]
        if self.loss_balancer is not None:
            main += list(self.loss_balancer.parameters())
        aux = [p for n, p in self.model.named_parameters() if n.endswith(".quantiles")]
        self.main_opt = torch.optim.Adam(main, lr=self.lr_main)
        self.sched = torch.optim.lr_scheduler.CosineAnnealingLR(self.main_opt, T_max=total_steps, eta_min=1e-8)
        self.aux_opt = torch.optim.Adam(aux, lr=self.lr_aux)

    def training_step(self, batch, with_metrics=False):
        loss, logs = self.rd_loss(batch, "train", with_metrics)
        self.main_opt.zero_grad()
        loss.backward()
        self.main_opt.step()
        aux = self.auxiliary_loss()
        logs["train/aux_loss"] = aux
        self.aux_opt.zero_grad()
        aux.backward()
        self.aux_opt.step()
        self.sched.step()
        if with_metrics:                                                      # after the optimiser steps, like mtc.py:468
            logs.update(self.average_metrics(*self._last, "train"))
        return loss, logs

    # ------------------------------------------------------------------ eval-time coding
    def update_bottleneck_values(self):
        self.model["compressor"].gaussian_conditional.update_scale_table(cai.get_scale_table())
        return self.model["compressor"].entropy_bottleneck.update()

    def compress(self, batch):
        ans = self.model["compressor"].compress(self.forward_input_heads(batch))
        return ans, sum(len(s) for part in ans["strings"] for s in part)

    def decompress(self, strings, shape):
        c = self.model["compressor"]
        z_hat = c.entropy_bottleneck.decompress(strings[1], shape)
        idx = c.gaussian_conditional.build_indexes(c.h_s(z_hat))
        y_hat = c.gaussian_conditional.decompress(strings[0], idx, z_hat.dtype)
        return self.forward_output_heads(c.g_s(y_hat))


def synthetic_batch(tasks: Sequence[str], B: int, size: int = 256, seed: int = 21, device="cpu"):
    """SURVEY.md 8d synthetic CLEVR-shaped inputs (seed 21 like src/train.py:204)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for t in tasks:
        c = TASKS[t]["in_channels"]
        if t == "semantic":
            v = torch.randint(0, 17, (B, c, size, size), generator=g).float()
        elif t == "depth_euclidean":
            v = torch.rand(B, c, size, size, generator=g) * 4.1
        else:
            v = torch.rand(B, c, size, size, generator=g)
        out[t] = v.to(device)
    return out


def latent_bits_expected(lik: Dict[str, torch.Tensor]) -> float:
    return float(sum((-torch.log2(v)).sum() for v in lik.values()))


__all__ = ["ReferenceCompressor", "TASKS", "synthetic_batch", "encoder_head", "decoder_head",
           "upsampled_decoder_head", "UncertaintyWeights", "Identity", "math"]
