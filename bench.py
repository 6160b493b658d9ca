#!/usr/bin/env python
"""Benchmark of the rate path (BASELINE.json metric: images/s rate-path fwd+bwd, 256^2 CLEVR, 3 tasks).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference (oracle/)
  python bench.py --config C3                              # another BASELINE configuration as the headline

Headline workload = BASELINE.json configs[1] ("C2"): MultiTaskDisjointLatentCompressor (-m 3), tasks rgb +
depth_euclidean + normal, -l 128 -c 100, lmbda 1e-2, 64 images per GPU (weak scaling), synthetic data.  The other
BASELINE configurations (C1, C3, C4) are measured the same way in the same run and reported under `other_configs`;
C5 (the rANS sweep, batch 1 ... 1024) is the `rans` object and runs at every N.

  value : images/s of ONE PASS OF THE HOT PATH with inputs resident in HBM: every GDN/IGDN site of the model
          forward + backward, EntropyBottleneck forward + backward, GaussianConditional forward + backward (with the
          reference's y (B,M,1,1) x scales (B,M,4,4) broadcast), the per-task distortion terms and the fused RD-loss
          epilogue, plus (N > 1) the all-reduce of the rate-path parameter gradients.  Convolutions are NOT in this
          number (north_star: they stay on cuDNN and are counted only end to end).  Eval configurations (C4) run
          the forward half only.
  e2e   : images/s of the whole step through the public API (`compressor.training_step(batch)` /
          `validation_step(batch)`: cuDNN convs + this repo's kernels + both optimizers), with the batch copied from
          pinned host memory and the loss read back every step.

Both arms print the same `metric`, `unit` and `config`; in both, `value` is the rate path and `e2e` the whole step
(`--impl reference`: the oracle on the host cores, each step a bounded sample of `--cpu-batch` images).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALL4 = ("rgb", "depth_euclidean", "normal", "semantic")
# BASELINE.json configs[0..3] (/root/reference/src/train.py:89-120, README.md:50-57)
CONFIGS = {
    "C1": dict(model_type=1, tasks=("mono",), latent_channels=32, conv_channels=32, lmbda=1e-2, batch=64, mode="train",
               workload="C1: SingleTaskCompressor -m 1 -t mono -l 32 -c 32 --lmbda 1e-2, 256x256"),
    "C2": dict(model_type=3, tasks=("rgb", "depth_euclidean", "normal"), latent_channels=128, conv_channels=100,
               lmbda=1e-2, batch=64, mode="train",
               workload="C2: MultiTaskDisjointLatentCompressor -m 3 -t rgb depth_euclidean normal -l 128 -c 100 --lmbda 1e-2, 256x256"),
    "C3": dict(model_type=2, tasks=ALL4, latent_channels=192, conv_channels=128, lmbda=1e-2, batch=64, mode="train",
               workload="C3: MultiTaskMixedLatentCompressor -m 2 -t rgb depth_euclidean normal semantic -l 192 -c 128 --lmbda 1e-2, 256x256, batch 64"),
    "C4": dict(model_type=4, tasks=ALL4, latent_channels=192, conv_channels=128, lmbda=1e-2, batch=256, mode="eval",
               workload="C4: MultiTaskSharedLatentCompressor -m 4 -t rgb depth_euclidean normal semantic -l 192 -c 128, "
                        "eval bpp / likelihood pass, 256x256, batch 256"),
}
METRIC = "images/s rate-path fwd+bwd (256x256 CLEVR-shaped, 3 tasks)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="headline configuration")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = the configuration's own batch)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per step of the CPU legs (a bounded sample)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-rans", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-likelihood", action="store_true", help="skip the EB / GC kernels at the declared roofline shape")
    ap.add_argument("--no-graph", action="store_true", help="launch the rate-path step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--serial-heads", action="store_true",
                    help="run the task heads' GDN sites one after the other on one stream (default: one stream per task head)")
    ap.add_argument("--precision", default="auto", help="GDN contraction: auto | fp32 | tf32 | 3xtf32")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "channels_last"],
                    help="memory format of the end-to-end step (the rate-path `value` always runs the NCHW kernels)")
    return ap.parse_args()


def config_dict(name, B, world, args):
    """The `config` object of the JSON line: a function of the command line only, so that both arms print the same."""
    return {"workload": CONFIGS[name]["workload"], "images_per_gpu": B, "global_batch": B * world,
            "parallelism": f"dp{world}", "gdn_precision": args.precision,
            "launch": "eager" if args.no_graph else "CUDA graph replay of one captured step",
            "l2_policy": "inputs larger than L2 (GBs of GDN activations per step vs 126 MB L2)"}


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- rate path
class RatePathHarness:
    """The hot path of one step with its inputs resident in HBM (no convolutions).

    Shapes are read off the real model: a single B=1 forward with hooks records the input shape of every GDN /
    IGDN site and of the two entropy models.  mode "train": forward + backward of everything; "eval": forward."""

    def __init__(self, mm, model, tasks, batch, device, torch, mode="train", concurrent_heads=True):
        self.mm, self.torch, self.model, self.B, self.mode, self.tasks = mm, torch, model, batch, mode, tasks
        train = mode == "train"
        sites, shapes = [], {}
        hooks = []
        for mod in model.modules():
            if isinstance(mod, mm.GDN):
                hooks.append(mod.register_forward_hook(lambda m, i, o: sites.append((m, tuple(i[0].shape)))))
        c = model.model["compressor"]
        hooks.append(c.entropy_bottleneck.register_forward_hook(lambda m, i, o: shapes.__setitem__("z", tuple(i[0].shape))))
        hooks.append(c.gaussian_conditional.register_forward_hook(
            lambda m, i, o: shapes.update(y=tuple(i[0].shape), s=tuple(i[1].shape))))
        with torch.no_grad():
            x_hats, _ = model(mm.synthetic_batch(tasks, 1, device=device))
        for h in hooks:
            h.remove()
        g = torch.Generator(device=device).manual_seed(21)
        rnd = lambda *s: torch.randn(*s, device=device, generator=g)  # noqa: E731
        # The T task heads are independent networks (mtc.py:109-177): their GDN sites may run concurrently, one CUDA
        # stream per head (the small layers launch 32-128 CTAs and leave most of the 148 SMs idle when serialised).
        # The step keeps the model's dependency structure: input heads (concurrent) -> backbone (alone) -> output
        # heads (concurrent); the backward of a site runs right after its forward, in the same phase.
        owner, phase_of = {}, {}
        for ph, key in ((0, "input_heads"), (2, "output_heads")):
            if key in model.model:
                for ti, head in enumerate(model.model[key]):
                    for m_ in head.modules():
                        owner[id(m_)] = ti
                        phase_of[id(m_)] = ph
        self.n_streams = (max(owner.values()) + 1) if (owner and concurrent_heads and len(tasks) > 1) else 1
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.n_streams)] if self.n_streams > 1 else []
        self.sites = []
        for mod, shp in sites:
            shp = (batch,) + shp[1:]
            x = rnd(*shp)
            self.sites.append((mod, x.requires_grad_(True) if train else x, rnd(*shp) if train else None,
                               owner.get(id(mod), 0) % self.n_streams, phase_of.get(id(mod), 1)))
        self.gdn_elems_per_image = sum(int(s_[1][0].numel()) for s_ in self.sites)
        zs, ys, ss = [(batch,) + shapes[k][1:] for k in ("z", "y", "s")]
        self.z = torch.distributions.Laplace(0.0, 2.0).sample(zs).to(device)
        self.scales = torch.exp(torch.empty(ss, device=device).uniform_(-3.0, 4.16, generator=g))
        self.y = rnd(*ys) * 3
        self.x = mm.synthetic_batch(tasks, batch, device=device, seed=22)
        self.x_hat = {}
        for t in tasks:
            shp = (batch,) + tuple(x_hats[t].shape[1:])
            self.x_hat[t] = (self.x[t] + 0.1 * rnd(*shp)) if shp == tuple(self.x[t].shape) else rnd(*shp)
        if train:
            for t_ in [self.z, self.scales, self.y] + list(self.x_hat.values()):
                t_.requires_grad_(True)
        self.eb, self.gc = c.entropy_bottleneck, c.gaussian_conditional
        # gradients of the rate-path parameters live in one flat bucket (exchanged when N > 1)
        self.eb_params = [p for n, p in self.eb.named_parameters() if n != "quantiles"]
        self.lv = list(model.loss_balancer.parameters())
        self.params = [p for s_ in self.sites for p in (s_[0].beta, s_[0].gamma)] + self.eb_params + self.lv
        self.bucket = mm.FlatGradBucket(self.params) if train else None
        self.loss_inputs = [self.z, self.y, self.scales] + list(self.x_hat.values()) + self.eb_params + self.lv

    def step(self, dist=None, world=1):
        torch, mm = self.torch, self.mm
        train = self.mode == "train"

        def site(mod, x, g):                  # K4: one GDN / IGDN site, forward (+ backward)
            y = mod(x)
            if train:
                _, gb, gg = torch.autograd.grad(y, [x, mod.beta, mod.gamma], g)
                mod.beta.grad.copy_(gb)
                mod.gamma.grad.copy_(gg)

        main = torch.cuda.current_stream()
        with torch.set_grad_enabled(train):
            for phase in (0, 1, 2):
                todo = [s_ for s_ in self.sites if s_[4] == phase]
                if phase == 1 or self.n_streams == 1:          # backbone (or serial mode): the caller's stream
                    for mod, x, g, _, _ in todo:
                        site(mod, x, g)
                    continue
                for st in self.streams:                        # fork: one stream per task head
                    st.wait_stream(main)
                for mod, x, g, sid, _ in todo:
                    with torch.cuda.stream(self.streams[sid]):
                        site(mod, x, g)
                for st in self.streams:                        # join before the next phase
                    main.wait_stream(st)
            self.eb.train(train), self.gc.train(train)
            mm.ops.noise_source.step()            # advance the device-side Philox stream (part of the captured graph)
            z_hat, z_lik = self.eb(self.z)        # K1 + K2
            y_hat, y_lik = self.gc(self.y, self.scales)  # K1 + K3
            lik = mm.compressors.LikelihoodDict(y=y_lik, z=z_lik)
            lik.log_sums = {"y": self.gc.last_log_likelihood_sums, "z": self.eb.last_log_likelihood_sums}
            loss, _ = self.model.rate_distortion_loss(self.x, self.x_hat, lik, "train" if train else "val")
            if train:
                grads = torch.autograd.grad(loss, self.loss_inputs)
                for p, gr in zip(self.eb_params + self.lv, grads[-(len(self.eb_params) + len(self.lv)):]):
                    p.grad.copy_(gr)
                if world > 1:                     # the path's one exchange step: rate-path parameter gradients
                    dist.all_reduce(self.bucket.flat)
        return loss

    def layer_table(self, precision):
        """Unique GDN layer shapes of the step with the kernel family each one takes (diagnostic entry points)."""
        mm = self.mm
        L, prec = mm._lib.lib(), mm.ops.GDN_PRECISION[precision]
        fam_f = {0: "streaming (C<=4)", 1: "fp32 SIMT", 2: "tcgen05", 3: "tcgen05 + TMA in/out",
                 6: "tcgen05, gamma streamed in K chunks (129-256 ch)"}
        fam_b = {0: "streaming (C<=4)", 1: "fp32 SIMT", 2: "fused tcgen05", 3: "fused tcgen05 + TMA, pipelined",
                 4: "fused tcgen05 + TMA, streamed gamma", 5: "fused tcgen05 + TMA, streamed gamma + x prefetch",
                 6: "tcgen05 dx (gamma, gamma^T streamed) + split-K tcgen05 d-gamma (129-256 ch)"}
        rows = {}
        for mod, x, g, _, _ in self.sites:
            B, C = x.shape[:2]
            HW = x.numel() // (B * C)
            key = (bool(mod.inverse), C, tuple(x.shape[2:]))
            if key not in rows:
                vf = int(L.mmnc_gdn_forward_variant(x.data_ptr(), x.data_ptr(), B, C, HW, prec))
                vb = int(L.mmnc_gdn_backward_variant(x.data_ptr(), x.data_ptr(), B, C, HW, prec))
                rows[key] = {"layer": "%sGDN(%d) @ %dx%d" % ("I" if mod.inverse else "", C, x.shape[2], x.shape[3]),
                             "sites": 0, "elements_per_image": int(x[0].numel()), "forward": fam_f.get(vf, str(vf)),
                             "backward": fam_b.get(vb, str(vb)) if self.mode == "train" else None}
            rows[key]["sites"] += 1
        return sorted(rows.values(), key=lambda r: -r["elements_per_image"] * r["sites"])


def time_kernel(torch, fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def roofline_of_gdn(torch, mm, harness, peak_gbs, peak_src, precision):
    """Dominant kernel of the step = the fused GDN backward on the largest layer (first GDN of each input head; the
    backward launches are > 50 % of the step's GPU time, profiles/).  Both kernels are launched alone through the
    C ABI (effective beta / gamma precomputed) on alternating inputs larger than L2 and timed with CUDA events on the
    launching stream.  `traffic` is the ncu DRAM byte count of the same launch (profiles/)."""
    L = mm._lib.lib()
    prec = mm.ops.GDN_PRECISION[precision]
    train = harness.mode == "train"
    big = sorted(harness.sites, key=lambda s: -s[1].numel())[:2]
    eff = []
    with torch.no_grad():
        for mod, x, g, _, _ in big:
            eff.append((mod.beta_reparam(mod.beta).clone(), mod.gamma_reparam(mod.gamma).clone(), x.detach(),
                        g if g is not None else torch.randn_like(x), mod.inverse))
    x0 = eff[0][2]
    B, C = x0.shape[:2]
    HW = x0.numel() // (B * C)
    n = x0.numel()
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty_like(x0)
    dbeta, dgamma = torch.empty_like(eff[0][0]), torch.empty_like(eff[0][1])
    nbytes = int(L.mmnc_gdn_backward_workspace_bytes(B, C, HW, prec))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
    nfw = int(L.mmnc_gdn_forward_workspace_bytes(B, C, HW, prec))  # > 0 only for the 129 .. 256 channel kernels
    wsf = torch.empty(max(nfw, 16), dtype=torch.uint8, device=x0.device)
    idx = [0]

    def fwd():
        b, gm, x, _, inv = eff[idx[0] % len(eff)]
        idx[0] += 1
        mm._lib.check(L.mmnc_gdn_forward_ws(x.data_ptr(), B, C, HW, b.data_ptr(), gm.data_ptr(), int(inv), prec,
                                            y.data_ptr(), wsf.data_ptr(), nfw, st))

    def bwd():
        b, gm, x, g, inv = eff[idx[0] % len(eff)]
        idx[0] += 1
        mm._lib.check(L.mmnc_gdn_backward(x.data_ptr(), g.data_ptr(), B, C, HW, b.data_ptr(), gm.data_ptr(), int(inv),
                                          prec, y.data_ptr(), dbeta.data_ptr(), dgamma.data_ptr(), ws.data_ptr(),
                                          nbytes, st))

    t_f = time_kernel(torch, fwd, reps=10)
    t_b = time_kernel(torch, bwd, reps=10)
    variant = int(L.mmnc_gdn_backward_variant(x0.data_ptr(), eff[0][3].data_ptr(), B, C, HW, prec))
    shape = "GDN(%d) on %dx%d, batch %d" % (C, x0.shape[2], x0.shape[3], B)
    # DRAM bytes of exactly this launch from `ncu --set full` (profiles/r01_ncu_summary.md); other shapes: unknown
    ncu_traffic = {(64, 50, 65536): (2.480e9, 1.632e9)}.get((B, C, HW), (None, None))
    fam = {5: "TMA-fed, streamed gamma + x prefetch", 4: "TMA-fed, streamed gamma", 3: "TMA-fed pipelined",
           2: "first generation"}.get(variant, "variant %d" % variant)
    bwd_roof = {
        "bound": "hbm", "kernel": "gdn backward (fused tcgen05, %s), %s; includes its ~10 us partial-reduce launch" % (fam, shape),
        "achieved": 12.0 * n / t_b / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 12.0 * n / t_b / 1e9 / peak_gbs,
        "traffic": ncu_traffic[0], "peak_source": peak_src, "algorithmic_bytes_per_launch": 12 * n,
        "launch_ms": t_b * 1e3}
    fwd_roof = {
        "bound": "hbm", "kernel": "gdn forward (tcgen05), " + shape,
        "achieved": 8.0 * n / t_f / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 8.0 * n / t_f / 1e9 / peak_gbs,
        "traffic": ncu_traffic[1], "peak_source": peak_src, "algorithmic_bytes_per_launch": 8 * n, "launch_ms": t_f * 1e3}
    out = {"roofline": bwd_roof, "roofline_forward": fwd_roof} if train else {"roofline": fwd_roof}
    # the other layer shapes of the step that move more than L2 holds, same measurement (fraction of the same peak)
    others, seen = [], {tuple(x0.shape)}
    for mod, x, g, _, _ in sorted(harness.sites, key=lambda s_: -s_[1].numel()):
        shp = tuple(x.shape)
        if shp in seen or x.numel() * 4 < 96e6:
            continue
        seen.add(shp)
        with torch.no_grad():
            b_, gm_ = mod.beta_reparam(mod.beta).clone(), mod.gamma_reparam(mod.gamma).clone()
        xs_ = [x.detach(), torch.randn_like(x)]
        gq_ = g if g is not None else xs_[1]
        Bq, Cq = shp[:2]
        HWq = x.numel() // (Bq * Cq)
        yq, dbq, dgq = torch.empty_like(x), torch.empty_like(b_), torch.empty_like(gm_)
        nbq = int(L.mmnc_gdn_backward_workspace_bytes(Bq, Cq, HWq, prec))
        wsq = torch.empty(nbq, dtype=torch.uint8, device=x.device)
        nfq = int(L.mmnc_gdn_forward_workspace_bytes(Bq, Cq, HWq, prec))
        wfq = torch.empty(max(nfq, 16), dtype=torch.uint8, device=x.device)
        k = [0]

        def fq():
            k[0] += 1
            mm._lib.check(L.mmnc_gdn_forward_ws(xs_[k[0] & 1].data_ptr(), Bq, Cq, HWq, b_.data_ptr(), gm_.data_ptr(),
                                                int(mod.inverse), prec, yq.data_ptr(), wfq.data_ptr(), nfq, st))

        def bq():
            k[0] += 1
            mm._lib.check(L.mmnc_gdn_backward(xs_[k[0] & 1].data_ptr(), gq_.data_ptr(), Bq, Cq, HWq, b_.data_ptr(),
                                              gm_.data_ptr(), int(mod.inverse), prec, yq.data_ptr(), dbq.data_ptr(),
                                              dgq.data_ptr(), wsq.data_ptr(), nbq, st))

        tf_ = time_kernel(torch, fq, reps=10)
        row = {"layer": "%sGDN(%d) on %dx%d, batch %d" % ("I" if mod.inverse else "", Cq, shp[2], shp[3], Bq),
               "forward_frac": 8.0 * x.numel() / tf_ / 1e9 / peak_gbs, "forward_ms": tf_ * 1e3}
        if train:
            tb_ = time_kernel(torch, bq, reps=10)
            row.update({"backward_frac": 12.0 * x.numel() / tb_ / 1e9 / peak_gbs, "backward_ms": tb_ * 1e3,
                        "backward_variant": int(L.mmnc_gdn_backward_variant(x.data_ptr(), gq_.data_ptr(), Bq, Cq, HWq, prec))})
        others.append(row)
        del xs_, yq, wsq
    out["roofline_other_layers"] = others
    return out


def roofline_of_likelihood(torch, mm, device, peak_gbs):
    """K2 / K3 at the declared roofline shape (what a shape-consistent ScaleHyperprior yields at
    256^2, batch 1024): z (1024, 512, 4, 4), y and scales (1024, 192, 16, 16).  Algorithmic bytes: EB forward 12 B /
    element, backward 16; GC forward 16 B / element (no broadcast), backward 24.  Also the measured maximum relative
    error of the forward likelihoods against the fp32 and the float64 oracle on a slice (north_star bar: 1e-5)."""
    out = {}
    eb = mm.EntropyBottleneck(512).to(device).train()
    gc = mm.GaussianConditional(None).to(device).train()
    g = torch.Generator(device=device).manual_seed(3)
    z = (torch.randn(1024, 512, 4, 4, device=device, generator=g) * 3).requires_grad_(True)
    y = (torch.randn(1024, 192, 16, 16, device=device, generator=g) * 3).requires_grad_(True)
    sc = torch.exp(torch.empty(1024, 192, 16, 16, device=device).uniform_(-3, 4.16, generator=g)).requires_grad_(True)
    packed = eb.packed_parameters().detach().requires_grad_(True)
    med = eb._get_medians().detach().reshape(-1)
    f_eb = lambda: mm.ops.entropy_bottleneck_forward(z, packed, med, True, 1e-9, seed=1)  # noqa: E731
    f_gc = lambda: mm.ops.gaussian_conditional_forward(y, sc, None, True, 0.11, 1e-9, seed=1)  # noqa: E731
    with torch.no_grad():
        t_ef, t_gf = time_kernel(torch, f_eb), time_kernel(torch, f_gc)

    def eb_fb():
        o, l, s = f_eb()
        torch.autograd.grad(s.sum(), [z, packed])

    def gc_fb():
        o, l, s = f_gc()
        torch.autograd.grad(s.sum(), [y, sc])

    t_eb, t_gb = time_kernel(torch, eb_fb) - t_ef, time_kernel(torch, gc_fb) - t_gf
    nz, ny = z.numel(), y.numel()
    for name, t, nbytes in (("eb_forward", t_ef, 12 * nz), ("eb_backward", t_eb, 16 * nz),
                            ("gc_forward", t_gf, 16 * ny), ("gc_backward", t_gb, 24 * ny)):
        out[name] = {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": nbytes / t / 1e9 / peak_gbs, "launch_ms": t * 1e3, "algorithmic_bytes_per_launch": nbytes}
    out["shape"] = "z (1024, 512, 4, 4); y, scales (1024, 192, 16, 16); training mode (Philox noise in-kernel)"
    # measured error of the forward likelihoods (eval mode = deterministic quantisation) against the oracle
    try:
        from oracle import compressai_ref as R

        eb.eval(), gc.eval()
        reb = R.EntropyBottleneck(512).eval()
        reb.load_state_dict({k: v.cpu() for k, v in eb.state_dict().items()})
        rgc = R.GaussianConditional(None).eval()
        zs, ys, ss = z[:16].detach(), y[:4].detach(), sc[:4].detach()
        with torch.no_grad():
            lz, ly = eb(zs)[1].cpu().double(), gc(ys, ss)[1].cpu().double()
            lz32, ly32 = reb(zs.cpu())[1].double(), rgc(ys.cpu(), ss.cpu())[1].double()
            reb64 = R.EntropyBottleneck(512).double().eval()
            reb64.load_state_dict({k: (v.cpu().double() if v.is_floating_point() else v.cpu()) for k, v in eb.state_dict().items()})
            lz64 = reb64(zs.cpu().double())[1]
            ly64 = R.GaussianConditional(None).double().eval()(ys.cpu().double(), ss.cpu().double())[1]

        def mx(a, b):
            m = b > 1e-3
            return float(((a - b).abs() / b)[m].max())

        out["max_rel_error_bulk"] = {"eb_vs_fp32_oracle": mx(lz, lz32), "eb_vs_float64_oracle": mx(lz, lz64),
                                     "gc_vs_fp32_oracle": mx(ly, ly32), "gc_vs_float64_oracle": mx(ly, ly64),
                                     "fp32_oracle_vs_float64_eb": mx(lz32, lz64), "fp32_oracle_vs_float64_gc": mx(ly32, ly64),
                                     "note": "likelihoods > 1e-3; the kernels evaluate a cancellation-free form (DESIGN.md 3), so "
                                             "they sit closer to the exact value than the fp32 oracle does"}
    except Exception as ex:  # auxiliary: never lose the line
        out["max_rel_error_bulk"] = {"error": repr(ex)}
    return out


# ----------------------------------------------------------------------------------------------------- CPU legs
def _cpu_threads():
    import torch

    torch.set_num_threads(os.cpu_count())
    return torch.get_num_threads()


def cpu_rate_path_stepper(cfg, cpu_batch: int):
    """-> (step function, description): the oracle (torch CPU ops in CompressAI's op order) on the same rate-path
    workload (every GDN site, EB, GC, distortion, RD loss; forward + backward), bounded sample of `cpu_batch` images."""
    import torch

    from oracle import compressai_ref as R
    from oracle import reference_models as orm

    _cpu_threads()
    torch.manual_seed(21)
    tasks = cfg["tasks"]
    ref = orm.ReferenceCompressor(cfg["model_type"], tasks, cfg["latent_channels"], cfg["conv_channels"], lmbda=cfg["lmbda"])
    sites, shapes, hooks = [], {}, []
    for mod in ref.modules():
        if isinstance(mod, R.GDN):
            hooks.append(mod.register_forward_hook(lambda m, i, o: sites.append((m, tuple(i[0].shape)))))
    c = ref.model["compressor"]
    hooks.append(c.entropy_bottleneck.register_forward_hook(lambda m, i, o: shapes.__setitem__("z", tuple(i[0].shape))))
    hooks.append(c.gaussian_conditional.register_forward_hook(
        lambda m, i, o: shapes.update(y=tuple(i[0].shape), s=tuple(i[1].shape))))
    batch = orm.synthetic_batch(tasks, cpu_batch)
    with torch.no_grad():
        x_hats0, _ = ref(batch)
    for h in hooks:
        h.remove()
    train = cfg["mode"] == "train"
    data = [(m, torch.randn(*s).requires_grad_(train), torch.randn(*s)) for m, s in sites]
    z = (torch.randn(shapes["z"]) * 2).requires_grad_(train)
    y = (torch.randn(shapes["y"]) * 3).requires_grad_(train)
    sc = torch.exp(torch.empty(shapes["s"]).uniform_(-3, 4.16)).requires_grad_(train)
    x_hat = {t: (torch.randn_like(x_hats0[t]) * 0.1 + (batch[t] if batch[t].shape == x_hats0[t].shape else 0)
                 ).requires_grad_(train) for t in tasks}
    ref.train(train)

    def step():
        with torch.set_grad_enabled(train):
            for m, x, g in data:
                out = m(x)
                if train:
                    torch.autograd.grad(out, [x, m.beta, m.gamma], g)
            z_hat, z_lik = c.entropy_bottleneck(z)
            y_hat, y_lik = c.gaussian_conditional(y, sc)
            rec, _ = ref.multitask_reconstruction_loss(batch, x_hat)
            comp, _ = ref.multitask_compression_loss({"y": y_lik, "z": z_lik}, x_hat)
            loss = ref.lmbda * rec + comp
            if train:
                loss.backward()
        return float(loss.detach())

    return step, (f"rate-path pass ({'forward + backward' if train else 'forward'}) of the same workload at batch {cpu_batch} "
                  f"(oracle = CPU restatement of CompressAI 1.2.4, torch CPU ops, {torch.get_num_threads()} threads)")


def cpu_whole_step_stepper(cfg, cpu_batch: int, total_steps: int):
    """-> step function: the reference's whole training (or validation) step on the host cores."""
    import torch

    from oracle import reference_models as orm

    _cpu_threads()
    torch.manual_seed(21)
    tasks = cfg["tasks"]
    ref = orm.ReferenceCompressor(cfg["model_type"], tasks, cfg["latent_channels"], cfg["conv_channels"], lmbda=cfg["lmbda"])
    batch = orm.synthetic_batch(tasks, cpu_batch)
    if cfg["mode"] == "train":
        ref.train()
        ref.configure_optimizers(total_steps=max(1, total_steps))

        def step():
            loss, _ = ref.training_step(batch, with_metrics=True)  # PSNR + MS-SSIM of every task, like mtc.py:468
            return float(loss.detach())
    else:
        ref.eval()

        def step():
            with torch.no_grad():
                loss, _ = ref.rd_loss(batch, "val", with_metrics=True)
            return float(loss)
    return step


def timed(fn, n, budget_s=None):
    t0, k = time.perf_counter(), 0
    while k < n:
        fn()
        k += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    return (time.perf_counter() - t0) / k, k


def cpu_baseline_leg(cfg, cpu_batch, budget_s=20.0):
    step, what = cpu_rate_path_stepper(cfg, cpu_batch)
    step()
    dt, n = timed(step, 20, budget_s)
    out = {"value": cpu_batch / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"{n} timed passes: {what}, {dt * 1e3:.0f} ms per pass"}
    whole = cpu_whole_step_stepper(cfg, cpu_batch, 8)
    whole()
    dt2, n2 = timed(whole, 4, budget_s / 2)
    out["e2e_value"] = cpu_batch / dt2
    out["e2e_sample"] = (f"{n2} whole {'training' if cfg['mode'] == 'train' else 'validation'} steps at batch {cpu_batch} "
                         f"(conv heads + backbone + rate path{' + both Adam steps' if cfg['mode'] == 'train' else ''} + PSNR / MS-SSIM of every task), "
                         f"{dt2 * 1e3:.0f} ms per step")
    return out


def run_reference(args):
    """--impl reference: the reference's implementation of the path restated on the host cores (oracle/), same
    metric / unit / config as the b200 arm; `value` = rate path, `e2e` = whole step, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    Bc = args.cpu_batch
    rate, what = cpu_rate_path_stepper(cfg, Bc)
    for _ in range(args.warmup):
        rate()
    dt_rate, _ = timed(rate, args.steps)
    whole = cpu_whole_step_stepper(cfg, Bc, args.steps + args.warmup)
    for _ in range(args.warmup):
        whole()
    dt_whole, _ = timed(whole, args.steps)
    import torch

    v, e = Bc / dt_rate, Bc / dt_whole
    sample = (f"each step = {what}; e2e = the whole {'training' if cfg['mode'] == 'train' else 'validation'} step (conv heads "
              f"+ backbone + rate path + both Adam steps + PSNR / MS-SSIM of every task, mtc.py:468) at the same batch {Bc}: a bounded sample of the "
              f"{B}-image step of the b200 arm (CPU time per image is flat in the batch); parity unpinned")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_rate * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, B, world, args),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "threads": torch.get_num_threads(), "images_per_step": Bc},
        "e2e": {"value": e, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "ms_per_step": dt_whole * 1e3},
    }))


# ----------------------------------------------------------------------------------------------------- rANS leg (C5)
def rans_sweep(torch, mm, model, device, world, rank, dist, batches=(1, 8, 64, 256, 1024), cpu=True):
    """BASELINE configs[4]: compress / decompress throughput at the C2 latent shapes (y: 128, z: 300 symbols per
    image), batch 1 ... 1024 PER GPU (weak scaling: every rank codes its own images, no collective), strings returned
    to the host.  Per batch size: wall MB/s := symbols x 4 B / wall time of compress (build_indexes + both entropy
    models' encode + D2H of the strings) resp. decompress (H2D + both decodes), max over ranks; and the kernel-only
    rate (CUDA events around the device part).  CPU oracle timed on a bounded sample in both marshalling variants
    (SURVEY.md 8d): faithful = CompressAI's per-image Python loop with five .tolist(); lean = one C call per batch."""
    c = model.model["compressor"]
    eb, gc = c.entropy_bottleneck, c.gaussian_conditional
    M, N = c.M, c.N
    g = torch.Generator(device=device).manual_seed(5 + rank)
    nmax = max(batches)
    scales_all = torch.exp(torch.empty(nmax, M, 1, 1, device=device).uniform_(-3.0, 4.16, generator=g))
    y_all = torch.randn(nmax, M, 1, 1, device=device, generator=g) * scales_all
    y_all[torch.rand(y_all.shape, device=device, generator=g) < 0.01] *= 50.0   # 1 % outliers -> bypass coding
    z_all = torch.randn(nmax, N, 1, 1, device=device, generator=g) * 4

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows, ok_all, keep = [], True, None
    for B in batches:
        y, z, scales = y_all[:B], z_all[:B], scales_all[:B]

        def enc():
            idx = gc.build_indexes(scales)
            return gc.compress(y, idx), eb.compress(z), idx

        def dec(ys, zs, idx):
            return gc.decompress(ys, idx), eb.decompress(zs, (1, 1))

        ys, zs, idx = enc()
        dec(ys, zs, idx)
        reps = 5 if B >= 256 else 10
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            ys, zs, idx = enc()
        torch.cuda.synchronize()
        t_enc = maxr((time.perf_counter() - t0) / reps)
        t0 = time.perf_counter()
        for _ in range(reps):
            yh, zh = dec(ys, zs, idx)
        torch.cuda.synchronize()
        t_dec = maxr((time.perf_counter() - t0) / reps)
        ok_all &= bool(torch.equal(yh, torch.round(y)))
        # kernel-only: the device part of both encodes / decodes (no host copies), CUDA events
        ysym, zsym = gc.quantize(y, "symbols"), mm.ops.quantize_symbols(z, eb._get_medians().detach().reshape(1, -1, 1, 1))
        tg, te = gc._rans_tables(), eb._rans_tables()

        def k_enc():
            mm.ops.rans_encode_device(ysym, idx, 0, tg)
            mm.ops.rans_encode_device(zsym, None, 1, te)

        # the upload workspace is shared by every call: keep private copies for the timing loop
        yb = tuple(t_.clone() for t_ in mm.ops.rans_upload(ys, device))
        zb = tuple(t_.clone() for t_ in mm.ops.rans_upload(zs, device))

        def k_dec():
            mm.ops.rans_decode_device(yb[0], yb[1], yb[2], B, idx, 0, M, tg)
            mm.ops.rans_decode_device(zb[0], zb[1], zb[2], B, None, 1, N, te)

        tk_enc, tk_dec = maxr(time_kernel(torch, k_enc, reps)), maxr(time_kernel(torch, k_dec, reps))
        nsym = world * B * (M + N)
        rows.append({"images_per_gpu": B, "encode_MBps": nsym * 4 / t_enc / 1e6, "decode_MBps": nsym * 4 / t_dec / 1e6,
                     "encode_wall_ms": t_enc * 1e3, "decode_wall_ms": t_dec * 1e3,
                     "encode_kernel_Msym_s": nsym / tk_enc / 1e6, "decode_kernel_Msym_s": nsym / tk_dec / 1e6,
                     "encode_kernel_us": tk_enc * 1e6, "decode_kernel_us": tk_dec * 1e6,
                     "output_bytes_per_image": (sum(map(len, ys)) + sum(map(len, zs))) / B})
        if B == nmax:
            keep = (ys, zs, y, z, scales)
    out = {"symbols_per_image": M + N, "n_gpus": world, "roundtrip_ok": ok_all, "sweep": rows,
           "what": "wall: compress()/decompress() of the public modules incl. host copies of the strings; kernel: device part "
                   "only (CUDA events); MB/s := symbols x 4 B / time, whole job over all ranks, max over ranks"}
    # headline numbers of the leg = the largest batch
    out.update({"images": nmax * world, "encode_MBps": rows[-1]["encode_MBps"], "decode_MBps": rows[-1]["decode_MBps"],
                "encode_Msym_s": rows[-1]["encode_MBps"] / 4, "decode_Msym_s": rows[-1]["decode_MBps"] / 4})
    ys, zs, y, z, scales = keep
    gc.eval(), eb.eval()
    with torch.no_grad():
        est_bits = float(-torch.log2(gc(y, scales)[1]).sum() - torch.log2(eb(z)[1]).sum())
    out["bits_actual_over_estimated"] = 8 * (sum(map(len, ys)) + sum(map(len, zs))) / est_bits
    if cpu and rank == 0:
        from oracle import compressai_ref as R

        n_cpu = 8
        rgc, reb = R.GaussianConditional(None), R.EntropyBottleneck(N)
        rgc.update_scale_table(R.get_scale_table())
        reb.load_state_dict({k: v.cpu() for k, v in eb.state_dict().items()})
        yc, sc, zc = y[:n_cpu].cpu(), scales[:n_cpu].cpu(), z[:n_cpu].cpu()
        for mode in ("faithful", "lean"):
            rgc.marshalling = reb.marshalling = mode
            t0 = time.perf_counter()
            ic = rgc.build_indexes(sc)
            a, b = rgc.compress(yc, ic), reb.compress(zc)
            t1 = time.perf_counter()
            rgc.decompress(a, ic), reb.decompress(b, (1, 1))
            t2 = time.perf_counter()
            out[f"cpu_{mode}_encode_MBps"] = n_cpu * (M + N) * 4 / (t1 - t0) / 1e6
            out[f"cpu_{mode}_decode_MBps"] = n_cpu * (M + N) * 4 / (t2 - t1) / 1e6
            out["bit_exact_vs_cpu_oracle"] = bool(a == ys[:n_cpu] and b == zs[:n_cpu])
        out["cpu_sample"] = f"{n_cpu} images, 1 thread (the C coder is sequential), per-image cost independent of the batch"
    return out


# ----------------------------------------------------------------------------------------------------- one configuration
def measure_config(name, args, env, steps, warmup, headline):
    """Rate path (`value`) and whole step (`e2e`) of one BASELINE configuration.  -> dict"""
    torch, mm, dist = env["torch"], env["mm"], env["dist"]
    world, rank, device = env["world"], env["rank"], env["device"]
    cfg = CONFIGS[name]
    B = args.batch if (args.batch and headline) else cfg["batch"]
    train = cfg["mode"] == "train"
    torch.manual_seed(21)
    model = mm.build_compressor(cfg["model_type"], cfg["tasks"], cfg["latent_channels"], cfg["conv_channels"],
                                lmbda=cfg["lmbda"])
    for m in model.modules():
        if isinstance(m, mm.GDN):
            m.precision = args.precision
    model.update_bottleneck_values()  # CPU, before .to(device): the reference's order (src/compress.py:101-105)
    model.to(device)
    model.train(train)
    torch.cuda.reset_peak_memory_stats(device)
    dp = mm.DataParallel(model) if (world > 1 and train) else None
    if train:
        model.configure_optimizers(total_steps=10 * (steps + warmup))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # -------- value: the hot path with inputs resident in HBM
    harness = RatePathHarness(mm, model, cfg["tasks"], B, device, torch, mode=cfg["mode"],
                              concurrent_heads=not args.serial_heads)
    mm.ops.noise_source.enable_device_state(device, seed=21)  # Philox (seed, offset) on the device: graph-safe
    run_step = lambda: harness.step(dist, world)  # noqa: E731
    graph, launches_per_step = None, None
    if not args.no_graph:
        # The step is hundreds of small launches: without a graph the host cannot issue them as fast as the GPU
        # retires them.  Capture once, replay per step (fresh noise each replay through the device-side Philox state).
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                harness.step(dist, world)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        launches_per_step0 = mm.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            harness.step(dist, world)
        launches_per_step = mm.launch_count() - launches_per_step0
        run_step = graph.replay
    for _ in range(max(warmup, 3)):
        run_step()
    barrier()
    sampler = ClockSampler(env["local"]) if (headline and rank == 0) else None
    if sampler:
        sampler.start()
    launches0 = mm.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run_step()
    e1.record()
    barrier()
    t_rate = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
    launches = (launches_per_step * steps) if graph is not None else (mm.launch_count() - launches0)
    mm.ops.noise_source.disable_device_state()
    del graph
    res = {"config": name, "workload": cfg["workload"], "mode": cfg["mode"], "images_per_gpu": B,
           "value": world * B / t_rate, "unit": "images/s", "ms_per_step": t_rate * 1e3, "steps": steps,
           "gpu_launches": int(launches), "gdn_elements_per_image": harness.gdn_elems_per_image,
           "head_streams": harness.n_streams}
    per_elem = 20.0 if train else 8.0  # GDN: 8 B forward + 12 B backward per element; entropy models + distortion < 1 %
    step_bytes = per_elem * harness.gdn_elems_per_image * B
    res["step_roofline"] = {"bound": "hbm", "achieved": step_bytes / t_rate / 1e9, "peak": env["peak"], "unit": "GB/s",
                            "frac": step_bytes / t_rate / 1e9 / env["peak"], "algorithmic_bytes_per_step": step_bytes,
                            "peak_source": env["peak_src"]}
    res["gdn_layers"] = harness.layer_table(args.precision)

    # -------- e2e: public API, host batch, H2D + D2H inside the timed region
    if not args.no_e2e:
        if harness.bucket is not None:
            del harness.bucket
            harness.bucket = None
        if args.layout == "channels_last":
            model.use_channels_last()
        host = mm.synthetic_batch(cfg["tasks"], B, seed=21 + rank, pin_memory=True)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        if dp is not None:  # re-home the gradients of the whole model into DataParallel's bucket
            dp.rebuild_bucket()
        # Input pipeline = pinned host batch -> device on a copy stream, double-buffered the way a data loader with
        # pin_memory + non_blocking prefetch does it: the H2D of the NEXT step's batch is issued right after this
        # step's kernels, so one full batch crosses PCIe inside every timed step but overlaps the compute.
        copy_stream = torch.cuda.Stream(device=device)

        def upload():
            with torch.cuda.stream(copy_stream):
                return {k: v.to(device, non_blocking=True) for k, v in host.items()}

        pending = [upload()]

        def e2e_step():
            cur = torch.cuda.current_stream()
            cur.wait_stream(copy_stream)
            dev_batch = pending[0]
            for v in dev_batch.values():
                v.record_stream(cur)
            loss = model.training_step(dev_batch) if train else model.validation_step(dev_batch)
            pending[0] = upload()
            return float(loss.item())  # D2H read of the step's result

        n_e2e = steps if headline else min(steps, 5)
        for _ in range(max(min(warmup, 5), 3)):
            e2e_step()
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            last = e2e_step()
        e1.record()
        barrier()
        t_e2e = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)) / n_e2e
        res["e2e"] = {"value": world * B / t_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                      "ms_per_step": t_e2e * 1e3, "steps": n_e2e, "last_loss": last,
                      "peak_memory_GB": torch.cuda.max_memory_allocated(device) / 1e9, "layout": args.layout,
                      "what": ("compressor.training_step(batch)" if train else "compressor.validation_step(batch)") +
                              " incl. PSNR + MS-SSIM of every task (the reference logs both on every step, mtc.py:468)" +
                              ": H2D of one pinned batch per step (copy stream, prefetched one step ahead), cuDNN convs (TF32 "
                              "allowed, torch default) + mmnc kernels" +
                              (", backward, gradient all-reduce (N>1), both Adam steps" if train else "") + ", loss.item()"}
    if sampler:
        res["clocks"] = sampler.stop()
    return res, model, harness


# ----------------------------------------------------------------------------------------------------- main arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import mmnc_b200 as mm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True  # let cuDNN pick its conv algorithms for the fixed shapes (convs are out of scope)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    env = dict(torch=torch, mm=mm, dist=dist, world=world, rank=rank, local=local, device=device, peak=peak, peak_src=src)

    head, model, harness = measure_config(args.config, args, env, args.steps, args.warmup, headline=True)
    cfg = CONFIGS[args.config]
    B = head["images_per_gpu"]

    # -------- roofline of the dominant kernel, likelihood kernels, rANS sweep, CPU baseline (outside the timed regions)
    extra = {}
    if rank == 0:
        extra.update(roofline_of_gdn(torch, mm, harness, peak, src, args.precision))
        if not args.no_likelihood:
            try:
                extra["roofline_likelihood"] = roofline_of_likelihood(torch, mm, device, peak)
            except Exception as ex:  # auxiliary: never lose the main line
                extra["roofline_likelihood"] = {"error": repr(ex)}
    if not args.no_rans:
        try:
            if args.config == "C2":
                rmodel = model
            else:
                c2 = CONFIGS["C2"]
                rmodel = mm.build_compressor(c2["model_type"], c2["tasks"], c2["latent_channels"], c2["conv_channels"])
                rmodel.update_bottleneck_values()
                rmodel.to(device)
            r = rans_sweep(torch, mm, rmodel, device, world, rank, dist, cpu=(world == 1 and not args.no_cpu))
            if rank == 0:
                extra["rans"] = r
        except Exception as ex:
            if world > 1:
                raise
            extra["rans"] = {"error": repr(ex)}
    del harness, model
    torch.cuda.empty_cache()

    others = []
    if world == 1 and not args.no_others:
        for name in sorted(CONFIGS):
            if name == args.config:
                continue
            try:
                res, m_, h_ = measure_config(name, args, env, min(args.steps, 5), 3, headline=False)
                del m_, h_
            except Exception as ex:
                res = {"config": name, "workload": CONFIGS[name]["workload"], "error": repr(ex)}
            torch.cuda.empty_cache()
            others.append(res)
    if rank == 0 and world == 1 and not args.no_cpu:
        extra["cpu_baseline"] = cpu_baseline_leg(cfg, args.cpu_batch)
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.config, B, world, args),
            "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks"), "e2e": head.get("e2e"),
            "detail": {"mode": head["mode"], "gdn_elements_per_image": head["gdn_elements_per_image"],
                       "head_streams": head["head_streams"], "gdn_layers": head["gdn_layers"]},
            "step_roofline": head["step_roofline"],
        }
        line.update(extra)
        if "cpu_baseline" not in line:
            line["cpu_baseline"] = None
        else:
            cb = line["cpu_baseline"]
            line["speedup_vs_cpu"] = {
                "rate_path": head["value"] / cb["value"],
                "e2e": (head["e2e"]["value"] / cb["e2e_value"]) if head.get("e2e") else None,
                "note": f"like for like: same workload, GPU at {B} images per step, CPU on a bounded sample of "
                        f"{args.cpu_batch} images per step ({cb['cores']} cores)"}
        if others:
            line["other_configs"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without running NCCL / CUDA-graph destructors: tearing down a communicator that was captured in a
        # CUDA graph hung the 2-GPU run at exit.  Everything has been synchronised and printed by now.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
