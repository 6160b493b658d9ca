"""Host-side mirror of the reference's four multi-task compressors (same class names, constructor arguments and
method names), built on this package's kernels instead of CompressAI + ~150 torch launches per step.

  MultiTaskCompressor                 /root/reference/src/models/multi_task_compressor.py:27-549
  MultiTaskMixedLatentCompressor      /root/reference/src/models/mixed_latent.py:15-162          (-m 2)
  MultiTaskDisjointLatentCompressor   /root/reference/src/models/disjoint_latent.py:14-194       (-m 3)
  MultiTaskSharedLatentCompressor     /root/reference/src/models/shared_latent.py:9-162          (-m 4)
  SingleTaskCompressor                /root/reference/src/models/single_task_compressor.py:13-55 (-m 1)
  UncertaintyWeightingStrategy        /root/reference/src/loss_balancing.py:21-54

Lightning is replaced by plain `nn.Module` + explicit optimizers (the training harness is out of scope,
SURVEY.md section 2 row 8); `training_step` keeps the reference's two-optimizer order (mtc.py:448-466).
What is different by design: the rate term is computed from the per-channel sums of ln(likelihood) that the
likelihood kernels already reduced, and the whole scalar part of the RD loss (group bpp, uncertainty weighting,
lmbda * rec + comp, and all their gradients) is ONE launch (`mmnc_rd_epilogue`).
"""
from __future__ import annotations

import os

from typing import Dict, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import metrics as _metrics
from . import ops
from .layers import GDN, conv, deconv
from .models import ScaleHyperprior, get_scale_table

# /root/reference/src/datasets/task_configs.py:7-33
task_parameters = {
    "depth_euclidean": {"in_channels": 1, "out_channels": 1, "loss_function": "mse"},
    "rgb": {"in_channels": 3, "out_channels": 3, "loss_function": "mse"},
    "semantic": {"in_channels": 1, "out_channels": 17, "loss_function": "cross-entropy"},
    "normal": {"in_channels": 3, "out_channels": 3, "loss_function": "mse"},
    "mono": {"in_channels": 1, "out_channels": 1, "loss_function": "mse"},
}


# id(compressor) -> {key: value}: head streams and group tables (see MultiTaskCompressor._cache)
_RUNTIME_CACHE: Dict[int, Dict[tuple, tuple]] = {}


class DummyModule(nn.Module):
    """/root/reference/src/utils.py:56-61"""

    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


class NoWeightingStrategy(DummyModule):
    """/root/reference/src/loss_balancing.py:15-18"""


class UncertaintyWeightingStrategy(nn.Module):
    """exp(-s_t) L_t + s_t (zeroed where L_t == 0).  The arithmetic runs inside `mmnc_rd_epilogue`; the stand-alone
    `forward` keeps the reference's dict-in / dict-out behaviour for callers that use it directly."""

    def __init__(self, num_tasks: int):
        super().__init__()
        self.log_vars = nn.Parameter(torch.zeros(num_tasks))

    def forward(self, task_losses: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        losses = torch.stack(list(task_losses.values()))
        weighted = (torch.exp(-self.log_vars) * losses + self.log_vars) * (losses != 0.0)
        out = dict(task_losses)
        out.update(zip(out, weighted))
        return out


class LikelihoodDict(dict):
    """{"y": ..., "z": ...} like CompressAI's, plus `.log_sums` = per-channel sums of ln(likelihood)."""

    log_sums: Optional[Dict[str, torch.Tensor]] = None


class MultiTaskCompressor(nn.Module):
    def __init__(self, compressor_backbone_class: type, tasks: Tuple[str], input_channels: Tuple[int],
                 output_channels: Optional[Tuple[int]] = None, latent_channels: int = 128, conv_channels: int = 100,
                 lmbda: float = 1, learning_rate_main=1e-5, learning_rate_aux=1e-3, **kwargs):
        super().__init__()
        self.compressor_backbone_class = compressor_backbone_class or ScaleHyperprior
        self.tasks = tuple(tasks)
        self.n_tasks = len(self.tasks)
        self.input_channels = tuple(input_channels)
        self.output_channels = tuple(output_channels) if output_channels is not None else tuple(
            task_parameters[t]["out_channels"] for t in self.tasks)
        assert self.n_tasks == len(self.input_channels)
        self.latent_channels = latent_channels
        self.conv_channels = conv_channels
        self.lmbda = lmbda
        self.learning_rate_main = learning_rate_main
        self.learning_rate_aux = learning_rate_aux
        self.kwargs = kwargs
        self.model: nn.ModuleDict = self._build_model()
        self.loss_balancer = UncertaintyWeightingStrategy(self.n_tasks)
        # per-instance switch (default from MMNC_SERIAL_HEADS at construction time); see _run_heads
        self.concurrent_heads = os.environ.get("MMNC_SERIAL_HEADS", "0") != "1"
        self._optimizers = None
        # set by parallel.DataParallel: grad_zero() replaces optimizer.zero_grad (one memset of the flat gradient
        # bucket), grad_sync() is called between backward() and optimizer.step()
        self.grad_sync = None
        self.grad_zero = None
        # step metrics (mtc.py:92, 359-384, 468): the reference evaluates PSNR and MS-SSIM of every task on EVERY step.
        # So does this class: `train_metrics_every` = 1 (0 = never, n = every n-th training step).  PSNR comes from the
        # distortion sums and MS-SSIM is one fused launch per scale (csrc/ssim.cu): ~1.3 ms per step at batch 64, 3 tasks.
        # (f1) run the whole network channels-last: cuDNN's tensor-core convolutions are NHWC kernels, and with NCHW
        # tensors it converts around every one of them (~27 % of the training step's GPU time, profiles/); the GDN
        # kernels take NHWC natively for the large layers.  Switch with use_channels_last().
        self.channels_last = False
        self.metrics = ("psnr", "ms-ssim")
        self.train_metrics_every = 1
        self._train_steps = 0

    def get_model_name(self):
        return self.__class__.__name__

    def _get_number_of_pixels(self, x_hats: Dict[str, torch.Tensor], task: str) -> int:
        B, _, H, W = x_hats[task].shape
        return B * H * W

    # ------------------------------------------------------------------ topology (mtc.py:109-193)
    def _build_heads(self, input_channels: Union[Sequence[int], int],
                     output_channels_per_head: Union[Sequence[int], int], is_deconv=False) -> nn.ModuleList:
        T = self.n_tasks
        cin = [input_channels] * T if isinstance(input_channels, int) else list(input_channels)
        cout = [output_channels_per_head] * T if isinstance(output_channels_per_head, int) else list(
            output_channels_per_head)
        assert len(cin) == T and len(cout) == T
        heads = []
        for ic, oc in zip(cin, cout):
            if is_deconv:
                m = ic // 2
                seq = [deconv(ic, m), GDN(m, inverse=True), conv(m, m, kernel_size=3, stride=1), GDN(m, inverse=True),
                       deconv(m, m), GDN(m, inverse=True), conv(m, m, kernel_size=3, stride=1), GDN(m, inverse=True),
                       deconv(m, oc), GDN(oc, inverse=True), deconv(oc, oc), GDN(oc, inverse=True),
                       conv(oc, oc, kernel_size=3, stride=1)]
            else:
                m = oc // 2
                seq = [conv(ic, m, kernel_size=3, stride=1), GDN(m), conv(m, oc), GDN(oc)]
                for _ in range(4):
                    seq += [conv(oc, oc), GDN(oc)]
            heads.append(nn.Sequential(*seq))
        return nn.ModuleList(heads)

    def _build_compression_backbone(self, input_channels: int, latent_channels: int) -> nn.Module:
        model = self.compressor_backbone_class(N=input_channels, M=latent_channels, **self.kwargs)
        model.g_a[0] = conv(input_channels, input_channels)
        model.g_s[-1] = deconv(input_channels, input_channels)
        return model

    def _build_model(self) -> nn.ModuleDict:
        raise NotImplementedError()

    # ------------------------------------------------------------------ forward (mtc.py:200-221, 491-505)
    # The T task heads are independent networks (mtc.py:109-177, 209-221): they run concurrently, one CUDA stream per
    # head.  Their deep layers launch a few dozen CTAs each and would leave most of the 148 SMs idle one after the
    # other (rate-path step 6.8 -> 5.9 ms, bench.py).  Autograd replays every backward op on its forward stream and
    # synchronises the streams itself; outputs that cross back to the caller's stream are recorded there so that
    # the caching allocator does not recycle them early.  `concurrent_heads = False` (or MMNC_SERIAL_HEADS=1) restores
    # the reference's one-after-the-other order; the results are identical either way.
    # CUDA streams and the small per-device lookup tables live in a module-level registry keyed by the compressor's
    # id, NOT on the nn.Module: `copy.deepcopy(model)` / `torch.save(model)` must not meet stream handles.
    @property
    def _cache(self) -> Dict[tuple, tuple]:
        return _RUNTIME_CACHE.setdefault(id(self), {})

    def __del__(self):
        _RUNTIME_CACHE.pop(id(self), None)

    def _run_heads(self, fns):
        """fns: one zero-argument callable per head, each returning a tensor.  -> list of results."""
        first = next(self.parameters(), None)
        if (not self.concurrent_heads or len(fns) < 2 or first is None or not first.is_cuda):
            return [f() for f in fns]
        dev = first.device
        main = torch.cuda.current_stream(dev)
        key = ("head_streams", str(dev))
        if key not in self._cache or len(self._cache[key]) < len(fns):
            self._cache[key] = [torch.cuda.Stream(device=dev) for _ in fns]
        streams = self._cache[key][: len(fns)]
        outs = []
        for f, st in zip(fns, streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                outs.append(f())
        for o, st in zip(outs, streams):
            main.wait_stream(st)
            o.record_stream(main)
        return outs

    def use_channels_last(self, enabled: bool = True):
        """Stores the convolution weights channels-last and feeds the heads channels-last inputs (same numbers)."""
        self.channels_last = bool(enabled)
        self.to(memory_format=torch.channels_last if enabled else torch.contiguous_format)
        return self

    def _as_model_format(self, batch):
        if not self.channels_last:
            return batch
        return {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in batch.items()}

    def forward_input_heads(self, batch) -> torch.Tensor:
        batch = self._as_model_format(batch)
        heads = self.model["input_heads"]
        return torch.concat(self._run_heads([(lambda i=i, t=t: heads[i](batch[t])) for i, t in enumerate(self.tasks)]),
                            dim=1)

    def _head_input(self, stacked_latent_values, i: int, task: str) -> torch.Tensor:
        """What output head i reads of the decoded latent (all of it, its channel group, or group + shared group)."""
        raise NotImplementedError()

    def forward_output_heads(self, stacked_latent_values, tasks: Optional[Sequence[str]] = None):
        """-> {task: x_hat}.  `tasks` (an extension, used by the selective decoder) restricts the work to a subset."""
        heads = self.model["output_heads"]
        sel = [(i, t) for i, t in enumerate(self.tasks) if tasks is None or t in tasks]
        outs = self._run_heads([(lambda i=i, t=t: heads[i](self._head_input(stacked_latent_values, i, t)))
                                for i, t in sel])
        return dict(zip([t for _, t in sel], outs))

    def forward(self, batch):
        out = self.model["compressor"](self.forward_input_heads(batch))
        lik = LikelihoodDict(out["likelihoods"])
        lik.log_sums = out.get("log_likelihood_sums")
        return self.forward_output_heads(out["x_hat"]), lik

    # ------------------------------------------------------------------ rate groups (a13)
    def _rate_groups(self):
        """-> (channel -> group id list (M, -1 = no rate term), [task index normalising each group], group names)."""
        raise NotImplementedError()

    def _group_tables(self, x_hats, device):
        px = tuple(self._get_number_of_pixels(x_hats, t) for t in self.tasks)
        key = (px, str(device))
        if key not in self._cache:
            chan, norm_task, names = self._rate_groups()
            inv = torch.tensor([1.0 / px[t] for t in norm_task], dtype=torch.float32, device=device)
            w = torch.full((len(norm_task),), 1.0 / self.n_tasks, dtype=torch.float32, device=device)
            self._cache[key] = (torch.tensor(chan, dtype=torch.int32, device=device), inv, w, 1.0 / px[0], names)
        return self._cache[key]

    def _log_sums(self, likelihoods):
        sums = getattr(likelihoods, "log_sums", None)
        if sums is None or sums.get("y") is None or sums.get("z") is None:
            sums = {k: ops.channel_log_likelihood_sums(v) for k, v in likelihoods.items()}
        return sums

    # ------------------------------------------------------------------ distortion (mtc.py:223-276)
    def reconstruction_loss(self, x_hat, x, loss_type: str = "mse") -> torch.Tensor:
        if loss_type in ("mse", "l1"):
            return ops.distortion(x_hat, x, loss_type)
        if loss_type == "cross-entropy":
            return F.cross_entropy(input=x_hat, target=x.squeeze(1).long(), reduction="mean")
        if loss_type == "ms-ssim":
            raise NotImplementedError("ms-ssim not implemented yet")
        raise NotImplementedError("reconstruction_loss_type should be one of [mse, ms-ssim]")

    def _task_losses(self, x, x_hats, log_dir, logs):
        vals = []
        for task in self.tasks:
            name = task_parameters[task]["loss_function"]
            vals.append(self.reconstruction_loss(x_hat=x_hats[task], x=x[task], loss_type=name))
            logs[f"{log_dir}/{task}/{name}"] = vals[-1].detach()
        return torch.stack(vals)

    def _log_vars(self):
        return getattr(self.loss_balancer, "log_vars", None)

    # ------------------------------------------------------------------ fused RD loss (a6 + a7 + mtc.py:437)
    def rate_distortion_loss(self, batch, x_hats, likelihoods, log_dir: str):
        """loss = lmbda * sum_t w_t(L_t) + comp, with every scalar the reference logs; one epilogue launch."""
        logs: Dict[str, torch.Tensor] = {}
        task_losses = self._task_losses(batch, x_hats, log_dir, logs)
        sums = self._log_sums(likelihoods)
        chan, inv, w, z_inv, names = self._group_tables(x_hats, task_losses.device)
        loss, s = ops.rd_epilogue(sums["y"], sums["z"], task_losses, self._log_vars(), chan, inv, w, z_inv,
                                  1.0 / self.n_tasks, self.lmbda)
        G = inv.numel()
        logs[f"{log_dir}/rec_loss"], logs[f"{log_dir}/compression_loss"], logs[f"{log_dir}/loss"] = s[1], s[2], s[0]
        self._compression_logs(logs, log_dir, names, s[4:4 + G], s[3])
        if self._log_vars() is not None:
            for i, task in enumerate(self.tasks):
                logs[f"uncertainty-weight/{task}"] = self.loss_balancer.log_vars[i].detach()
        return loss, logs

    def _compression_logs(self, logs, log_dir, names, group_bpp, z_bpp):
        for i, name in enumerate(names):
            logs[f"{log_dir}/{name}/compression_loss"] = group_bpp[i] + z_bpp

    # reference-shaped entry points (each returns (value, logs)); same kernels, partial epilogues
    def multitask_reconstruction_loss(self, x, x_hats, log_dir: str):
        logs: Dict[str, torch.Tensor] = {}
        task_losses = self._task_losses(x, x_hats, log_dir, logs)
        dev = task_losses.device
        empty = torch.zeros(0, dtype=torch.float32, device=dev)
        rec, _ = ops.rd_epilogue(empty, empty, task_losses, self._log_vars(),
                                 torch.zeros(0, dtype=torch.int32, device=dev), empty, empty, 0.0, 0.0, 1.0)
        if self._log_vars() is not None:
            for i, task in enumerate(self.tasks):
                logs[f"uncertainty-weight/{task}"] = self.loss_balancer.log_vars[i].detach()
        return rec, logs

    def _bits_per_pixel(self, likelihoods: torch.Tensor, num_pixels) -> torch.Tensor:
        return ops.channel_log_likelihood_sums(likelihoods).sum() / (-torch.log(torch.tensor(2.0))) / num_pixels

    def multitask_compression_loss(self, all_likelihoods, x_hats, log_dir: str):
        logs: Dict[str, torch.Tensor] = {}
        sums = self._log_sums(all_likelihoods)
        dev = sums["y"].device
        chan, inv, w, z_inv, names = self._group_tables(x_hats, dev)
        loss, s = ops.rd_epilogue(sums["y"], sums["z"], torch.zeros(0, dtype=torch.float32, device=dev), None, chan, inv,
                                  w, z_inv, 1.0 / self.n_tasks, 0.0)
        self._compression_logs(logs, log_dir, names, s[4:4 + inv.numel()], s[3])
        return loss, logs

    # ------------------------------------------------------------------ metrics (mtc.py:359-384)
    def average_metrics(self, x, x_hats, log_dir: str, task_losses=None) -> Dict[str, torch.Tensor]:
        """PSNR / MS-SSIM per task, same log names as the reference.  `task_losses` = {task: distortion term} lets the
        PSNR of the mse tasks reuse the sums the loss already reduced."""
        return _metrics.average_metrics(self.tasks, x, x_hats, log_dir, task_losses,
                                        with_ms_ssim="ms-ssim" in self.metrics)

    # ------------------------------------------------------------------ optimisation (mtc.py:386-418)
    def auxiliary_loss(self):
        return self.model["compressor"].entropy_bottleneck.loss()

    def get_main_parameters(self):
        return [p for n, p in self.model.named_parameters() if not n.endswith(".quantiles")]

    def get_auxiliary_parameters(self):
        return [p for n, p in self.model.named_parameters() if n.endswith(".quantiles")]

    def configure_optimizers(self, total_steps: int):
        """`total_steps` = the run's number of optimizer steps (the reference passes Lightning's
        `estimated_stepping_batches`, mtc.py:405-409): required, because a cosine schedule with a guessed horizon
        turns upwards again once the guess is passed."""
        if int(total_steps) < 1:
            raise ValueError("total_steps must be a positive number of optimizer steps")
        fused = next(self.parameters()).is_cuda
        main = torch.optim.Adam(self.get_main_parameters() + list(self.loss_balancer.parameters()),
                                lr=self.learning_rate_main, fused=fused)
        sch = torch.optim.lr_scheduler.CosineAnnealingLR(main, T_max=total_steps, eta_min=1e-8)
        aux = torch.optim.Adam(self.get_auxiliary_parameters(), lr=self.learning_rate_aux, fused=fused)
        self._optimizers = (main, aux, sch)
        return {"optimizer": main, "lr_scheduler": {"scheduler": sch}}, {"optimizer": aux}

    def optimizers(self):
        if self._optimizers is None:
            raise RuntimeError("call configure_optimizers(total_steps=...) before the first training_step")
        return self._optimizers[0], self._optimizers[1]

    def lr_schedulers(self):
        return self._optimizers[2]

    # ------------------------------------------------------------------ steps (mtc.py:420-483)
    def _step(self, batch, is_train: bool):
        log_dir = "train" if is_train else "val"
        batch = self._as_model_format(batch)  # once: the loss reads the inputs in the same layout as the outputs
        x_hats, likelihoods = self.forward(batch)
        loss, log_dict = self.rate_distortion_loss(batch, x_hats, likelihoods, log_dir)
        # Step metrics (mtc.py:468).  They depend only on the inputs and the reconstructions, so on a GPU they are issued
        # NOW, on their own stream, and run underneath the backward pass instead of after the optimiser steps.  `batch` and
        # `x_hats` stay referenced until the end of this function, where the main stream waits for the metric stream, so
        # nothing they read can be reused early.
        self._train_steps += 1 if is_train else 0
        every = self.train_metrics_every if is_train else 1
        want_metrics = bool(self.metrics) and bool(every) and (not is_train or self._train_steps % every == 0)
        metric_logs, metric_stream = None, None
        if want_metrics:
            reuse = {t: log_dict[f"{log_dir}/{t}/mse"] for t in self.tasks
                     if task_parameters[t]["loss_function"] == "mse" and f"{log_dir}/{t}/mse" in log_dict}
            if is_train and loss.is_cuda:
                key = ("metric_stream", str(loss.device))
                if key not in self._cache:
                    self._cache[key] = torch.cuda.Stream(device=loss.device)
                metric_stream = self._cache[key]
                metric_stream.wait_stream(torch.cuda.current_stream(loss.device))
                with torch.cuda.stream(metric_stream):
                    metric_logs = self.average_metrics(batch, x_hats, log_dir, reuse)
            else:
                metric_logs = self.average_metrics(batch, x_hats, log_dir, reuse)
        if is_train:
            main_opt, aux_opt = self.optimizers()
            if self.grad_zero is not None:
                self.grad_zero()
            else:
                main_opt.zero_grad(set_to_none=True)  # backward then assigns the gradients: no zeroing, no adds
            loss.backward()
            if self.grad_sync is not None:
                self.grad_sync()
            main_opt.step()
            aux_loss = self.auxiliary_loss()
            log_dict[f"{log_dir}/aux_loss"] = aux_loss.detach()
            aux_opt.zero_grad(set_to_none=True)
            aux_loss.backward()
            aux_opt.step()
            self.lr_schedulers().step()
        if metric_stream is not None:
            torch.cuda.current_stream(loss.device).wait_stream(metric_stream)
        if metric_logs is not None:
            log_dict.update(metric_logs)
        self.last_logs = log_dict
        return loss

    def training_step(self, batch, batch_idx: int = 0):
        return self._step(batch, is_train=True)

    @torch.no_grad()
    def validation_step(self, batch, batch_idx: int = 0):
        # Lightning calls model.eval() around validation (mtc.py:482-483 relies on it): quantisation must be round(),
        # not additive noise, whatever mode the caller left the module in.  The previous mode is restored.
        was_training = self.training
        self.eval()
        try:
            return self._step(batch, is_train=False)
        finally:
            self.train(was_training)

    # ------------------------------------------------------------------ eval-time coding (mtc.py:486-549)
    def update_bottleneck_values(self):
        self.model["compressor"].gaussian_conditional.update_scale_table(get_scale_table())
        return self.model["compressor"].entropy_bottleneck.update()

    @torch.no_grad()
    def compress(self, batch, print_info: bool = False):
        stacked_t = self.forward_input_heads(batch)
        ans = self.model["compressor"].compress(stacked_t)
        number_of_bytes = sum(len(s) for latents in ans["strings"] for s in latents)
        stacked_t_likelihoods = None
        if print_info:
            B, _, H, W = batch[self.tasks[0]].shape
            bpp = number_of_bytes * 8 / B / H / W / self.n_tasks
            print(f"Number of actual bytes in a string is: {number_of_bytes}, which gives a BPP = {bpp:.3f}")
            out = self.model["compressor"](stacked_t)
            stacked_t_likelihoods = LikelihoodDict(out["likelihoods"])
            stacked_t_likelihoods.log_sums = out.get("log_likelihood_sums")
            compression_loss, _ = self.multitask_compression_loss(stacked_t_likelihoods, x_hats=batch, log_dir="")
            print(f"Estimated BPP (compression loss) is: {compression_loss.item():.3f}")
        return ans, number_of_bytes, stacked_t_likelihoods

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        c = self.model["compressor"]
        z_hat = c.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat = c.h_s(z_hat)
        indexes = c.gaussian_conditional.build_indexes(scales_hat)
        y_hat = c.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        return self.forward_output_heads(c.g_s(y_hat))


    # ------------------------------------------------------------------ (f3) container with per-group streams
    def _coding_groups(self):
        """[(name, first channel, channel count)]: the rate groups as contiguous channel ranges; channels that belong
        to no group (the Disjoint model's orphans) feed no decoder and are not coded."""
        chan, _, names = self._rate_groups()
        groups = []
        for gi, name in enumerate(names):
            idx = [c for c, g in enumerate(chan) if g == gi]
            assert idx and idx == list(range(idx[0], idx[0] + len(idx))), "rate groups are contiguous channel ranges"
            groups.append((name, idx[0], len(idx)))
        return groups

    def _groups_for_tasks(self, tasks: Optional[Sequence[str]]):
        names = [g[0] for g in self._coding_groups()]
        if tasks is None or set(names) <= {"__all__"}:
            return names
        unknown = [t for t in tasks if t not in self.tasks]
        if unknown:
            raise KeyError(f"unknown task(s) {unknown}; this model codes {list(self.tasks)}")
        return [n for n in names if n in tasks or n == "shared"]

    @staticmethod
    def _coding_scales(y_shape, scales_hat: torch.Tensor) -> torch.Tensor:
        """The scales the y symbols are coded against.  Shape-consistent models: scales_hat itself.  The reference's
        256^2 geometry (y (B,M,1,1) against scales (B,M,4,4): `compress` raises there) has sixteen
        candidate scales per symbol; the container codes each symbol against their spatial MEAN - encoder and decoder
        derive it from the same z_hat, so the choice only affects the rate, never decodability."""
        if tuple(scales_hat.shape[-2:]) == tuple(y_shape):
            return scales_hat
        if all(d == 1 for d in y_shape):
            return scales_hat.mean(dim=(2, 3), keepdim=True)
        raise ValueError(f"cannot code y with spatial shape {tuple(y_shape)} against scales {tuple(scales_hat.shape)}")

    @torch.no_grad()
    def compress_to_container(self, batch):
        """Inputs -> `container.Container` (call `.to_bytes()` for the on-disk form): z once, y group by group."""
        from .container import Container

        c = self.model["compressor"]
        y = c.g_a(self.forward_input_heads(batch))
        z = c.h_a(torch.abs(y))
        eb, gc = c.entropy_bottleneck, c.gaussian_conditional
        z_strings = eb.compress(z)
        z_hat = eb.quantize(z, "dequantize", eb._get_medians().detach().reshape(1, -1, *([1] * (z.dim() - 2))))
        indexes = gc.build_indexes(self._coding_scales(y.shape[-2:], c.h_s(z_hat)))
        groups = self._coding_groups()
        y_strings = {name: gc.compress(y[:, a:a + n].contiguous(), indexes[:, a:a + n].contiguous())
                     for name, a, n in groups}
        kind = {"SingleTaskCompressor": 1, "MultiTaskMixedLatentCompressor": 2, "MultiTaskDisjointLatentCompressor": 3,
                "MultiTaskSharedLatentCompressor": 4}.get(type(self).__name__, 0)
        return Container(kind, c.M, c.N, z.shape[-2:], y.shape[-2:], groups, z_strings, y_strings)

    @torch.no_grad()
    def decompress_container(self, source, tasks: Optional[Sequence[str]] = None):
        """Container (or its bytes) -> {task: x_hat} for `tasks` (default: all).  Only z and the channel groups those
        tasks read are sliced out of the payload and decoded, and only their output heads run."""
        from .container import Container

        wanted = self._groups_for_tasks(tasks)
        cont = Container.from_bytes(source, groups=wanted) if isinstance(source, (bytes, bytearray, memoryview)) else source
        c = self.model["compressor"]
        if (cont.M, cont.N) != (c.M, c.N):
            raise ValueError(f"container was written by a model with M = {cont.M}, N = {cont.N}; this one has {c.M}, {c.N}")
        eb, gc = c.entropy_bottleneck, c.gaussian_conditional
        z_hat = eb.decompress(cont.z_strings, cont.z_shape)
        indexes = gc.build_indexes(self._coding_scales(cont.y_shape, c.h_s(z_hat)))
        y_hat = torch.zeros((cont.n_images, c.M) + tuple(cont.y_shape), dtype=z_hat.dtype, device=z_hat.device)
        table = {name: (a, n) for name, a, n in cont.groups}
        for name in wanted:
            a, n = table[name]
            y_hat[:, a:a + n] = gc.decompress(cont.y_strings[name], indexes[:, a:a + n].contiguous(), z_hat.dtype)
        return self.forward_output_heads(c.g_s(y_hat), tasks=tasks)


class MultiTaskMixedLatentCompressor(MultiTaskCompressor):
    """All tasks share all M latent channels; every output head sees the whole latent (mixed_latent.py)."""

    def _rate_groups(self):
        M = self.model["compressor"].M
        return [0] * M, [0], ["__all__"]

    def _compression_logs(self, logs, log_dir, names, group_bpp, z_bpp):
        for task in self.tasks:  # mixed_latent.py:108-110: every task reports the full y + z rate
            logs[f"{log_dir}/{task}/compression_loss"] = group_bpp[0] + z_bpp

    def _get_task_likelihoods(self, likelihoods, task):
        return likelihoods

    def _build_model(self) -> nn.ModuleDict:
        model = nn.ModuleDict()
        model["input_heads"] = self._build_heads(self.input_channels, self.conv_channels)
        total = self.conv_channels * self.n_tasks
        model["compressor"] = self._build_compression_backbone(total, self.latent_channels)
        model["output_heads"] = self._build_heads(total, self.output_channels, is_deconv=True)
        return model

    def _head_input(self, stacked_latent_values, i, task):
        return stacked_latent_values


class SingleTaskCompressor(MultiTaskMixedLatentCompressor):
    """One task, no loss weighting (single_task_compressor.py:55)."""

    def __init__(self, compressor_backbone_class, tasks, input_channels, latent_channels, conv_channels,
                 lmbda: float = 1, learning_rate_main=1e-5, learning_rate_aux=1e-3, output_channels=None, **kwargs):
        assert len(tasks) == 1
        super().__init__(compressor_backbone_class=compressor_backbone_class, tasks=tasks,
                         input_channels=input_channels, output_channels=output_channels,
                         conv_channels=conv_channels, latent_channels=latent_channels, lmbda=lmbda,
                         learning_rate_main=learning_rate_main, learning_rate_aux=learning_rate_aux, **kwargs)
        self.loss_balancer = NoWeightingStrategy()


class MultiTaskDisjointLatentCompressor(MultiTaskCompressor):
    """Latent channels sliced per task, g_s removed, four extra deconvs per output head (disjoint_latent.py)."""

    def __init__(self, compressor_backbone_class, tasks, input_channels, output_channels=None, latent_channels=128,
                 conv_channels=100, lmbda: float = 1, learning_rate_main=1e-5, learning_rate_aux=1e-3, **kwargs):
        self.latent_channels_per_task = latent_channels // len(tasks)
        super().__init__(compressor_backbone_class=compressor_backbone_class, tasks=tasks,
                         input_channels=input_channels, output_channels=output_channels,
                         conv_channels=conv_channels, latent_channels=latent_channels, lmbda=lmbda,
                         learning_rate_main=learning_rate_main, learning_rate_aux=learning_rate_aux, **kwargs)
        if self.latent_channels % self.n_tasks != 0:
            # disjoint_latent.py:68-75: cosmetic — the backbone was already built with the unrounded count, so
            # the trailing channels are coded but carry no rate term and feed no decoder
            self.latent_channels = self.latent_channels_per_task * self.n_tasks

    def _group_width(self) -> int:
        return self.latent_channels_per_task

    def _rate_groups(self):
        M, k, T = self.model["compressor"].M, self._group_width(), self.n_tasks
        chan = [(c // k) if c < k * T else -1 for c in range(M)]
        return chan, list(range(T)), list(self.tasks)

    def _get_task_channels(self, tensor: torch.Tensor, task: str) -> torch.Tensor:
        assert tensor.dim() == 4
        i, k = self.tasks.index(task), self._group_width()
        return tensor[:, i * k:(i + 1) * k, :, :]

    def _get_task_likelihoods(self, likelihoods, task):
        return self._get_task_channels(likelihoods["y"], task)

    def _build_heads(self, input_channels, output_channels_per_head, is_deconv=False) -> nn.ModuleList:
        if not is_deconv:
            return super()._build_heads(input_channels, output_channels_per_head, is_deconv)
        w = self.conv_channels // self.n_tasks
        tails = super()._build_heads(self.conv_channels, output_channels_per_head, is_deconv)
        heads = nn.ModuleList()
        for i in range(self.n_tasks):
            heads.append(nn.Sequential(deconv(input_channels, w), GDN(w, inverse=True), deconv(w, w),
                                       GDN(w, inverse=True), deconv(w, w), GDN(w, inverse=True),
                                       deconv(w, self.conv_channels), tails[i]))
        return heads

    def _head_width(self) -> int:
        return self.latent_channels_per_task

    def _build_model(self) -> nn.ModuleDict:
        model = nn.ModuleDict()
        model["input_heads"] = self._build_heads(self.input_channels, self.conv_channels)
        total = self.conv_channels * self.n_tasks
        model["compressor"] = self._build_compression_backbone(total, self.latent_channels)
        model["compressor"].g_s = DummyModule()
        model["output_heads"] = self._build_heads(self._head_width(), self.output_channels, is_deconv=True)
        return model

    def _head_input(self, stacked_latent_values, i, task):
        return self._get_task_channels(stacked_latent_values, task)


class MultiTaskSharedLatentCompressor(MultiTaskDisjointLatentCompressor):
    """T task-specific channel groups plus one shared group that every head also reads (shared_latent.py)."""

    def __init__(self, compressor_backbone_class, tasks, input_channels, output_channels=None, latent_channels=192,
                 conv_channels=128, lmbda: float = 1, learning_rate_main=1e-5, learning_rate_aux=1e-3, **kwargs):
        n = len(tasks)
        if latent_channels % (n + 1) != 0:
            latent_channels = latent_channels // (n + 1) * (n + 1)  # shared_latent.py:34-41
        self.task_specific_channels_n = latent_channels // (n + 1)
        super().__init__(compressor_backbone_class=compressor_backbone_class, tasks=tasks,
                         input_channels=input_channels, output_channels=output_channels,
                         conv_channels=conv_channels, latent_channels=latent_channels, lmbda=lmbda,
                         learning_rate_main=learning_rate_main, learning_rate_aux=learning_rate_aux, **kwargs)

    def _group_width(self) -> int:
        return self.task_specific_channels_n

    def _head_width(self) -> int:
        return self.task_specific_channels_n * 2

    def _rate_groups(self):
        M, k, T = self.model["compressor"].M, self._group_width(), self.n_tasks
        chan = [min(c // k, T) for c in range(M)]  # M == (T + 1) * k; the last group is the shared one
        return chan, list(range(T)) + [0], list(self.tasks) + ["shared"]

    def _shared_channels(self, tensor: torch.Tensor) -> torch.Tensor:
        return tensor[:, -self.task_specific_channels_n:, :, :]

    def _get_task_likelihoods(self, likelihoods, task):
        if task == "shared":
            return self._shared_channels(likelihoods["y"])
        return self._get_task_channels(likelihoods["y"], task)

    def _head_input(self, stacked_latent_values, i, task):
        B, _, H, W = stacked_latent_values.shape
        own = self._get_task_channels(stacked_latent_values, task)
        return torch.stack([own, self._shared_channels(stacked_latent_values)], dim=1).reshape((B, -1, H, W))


def build_compressor(model_type: int, tasks: Sequence[str], latent_channels: int, conv_channels: int,
                     lmbda: float = 1.0, **kw) -> MultiTaskCompressor:
    """The reference CLI's `-m {1,2,3,4} -t ... -l ... -c ... --lmbda ...` (/root/reference/src/train.py:89-120)."""
    cls = {1: SingleTaskCompressor, 2: MultiTaskMixedLatentCompressor, 3: MultiTaskDisjointLatentCompressor,
           4: MultiTaskSharedLatentCompressor}[int(model_type)]
    cin = tuple(task_parameters[t]["in_channels"] for t in tasks)
    cout = tuple(task_parameters[t]["out_channels"] for t in tasks)
    return cls(compressor_backbone_class=ScaleHyperprior, tasks=tuple(tasks), input_channels=cin,
               output_channels=cout, latent_channels=latent_channels, conv_channels=conv_channels, lmbda=lmbda, **kw)
