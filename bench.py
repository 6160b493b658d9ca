#!/usr/bin/env python
"""Benchmark of the rate path (BASELINE.json metric: images/s rate-path fwd+bwd, 256^2 CLEVR, 3 tasks).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference (oracle/)

Workload = BASELINE.json configs[1] ("C2"): MultiTaskDisjointLatentCompressor (-m 3), tasks rgb +
depth_euclidean + normal, -l 128 -c 100, lmbda 1e-2, 64 images per GPU (weak scaling), synthetic data.

  value : images/s of ONE PASS OF THE HOT PATH with inputs resident in HBM: every GDN/IGDN site of the model
          (48 calls, 18.5 M elements per image) forward + backward, EntropyBottleneck forward + backward,
          GaussianConditional forward + backward (with the reference's y (B,M,1,1) x scales (B,M,4,4) broadcast),
          the per-task distortion terms and the fused RD-loss epilogue, plus (N > 1) the all-reduce of the
          rate-path parameter gradients.  Convolutions are NOT in this number (north_star: they stay on cuDNN
          and are counted only end to end).
  e2e   : images/s of the whole training step through the public API (`compressor.training_step(batch)`:
          cuDNN convs + this repo's kernels + both optimizers), with the batch copied from pinned host memory
          and the loss read back every step.  This is the number to hold against `--impl reference`, which
          runs the same training step with the oracle modules on the host cores.
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

TASKS = ("rgb", "depth_euclidean", "normal")
MODEL = dict(model_type=3, latent_channels=128, conv_channels=100, lmbda=1e-2)
METRIC = "images/s rate-path fwd+bwd (256x256 CLEVR-shaped, 3 tasks)"
WORKLOAD = "C2: MultiTaskDisjointLatentCompressor -m 3 -t rgb depth_euclidean normal -l 128 -c 100 --lmbda 1e-2, 256x256"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--cpu-batch", type=int, default=1, help="images per step of the CPU legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-rans", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the rate-path step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--serial-heads", action="store_true",
                    help="run the task heads' GDN sites one after the other on one stream (default: one stream per task head)")
    ap.add_argument("--precision", default="auto", help="GDN contraction: auto | fp32 | tf32 | 3xtf32")
    return ap.parse_args()


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
    """The hot path of one training step with its inputs resident in HBM (no convolutions).

    Shapes are read off the real model: a single B=1 forward with hooks records the input shape of every GDN /
    IGDN site and of the two entropy models."""

    def __init__(self, mm, model, batch, device, torch, concurrent_heads=True):
        self.mm, self.torch, self.model, self.B = mm, torch, model, batch
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
            x_hats, _ = model(mm.synthetic_batch(TASKS, 1, device=device))
        for h in hooks:
            h.remove()
        g = torch.Generator(device=device).manual_seed(21)
        rnd = lambda *s: torch.randn(*s, device=device, generator=g)  # noqa: E731
        # The T task heads are independent networks (mtc.py:109-177): their GDN sites may run concurrently, one CUDA stream
        # per head (the small layers launch 32-128 CTAs and leave most of the 148 SMs idle when serialised).
        # The step keeps the model's dependency structure: input heads (concurrent) -> backbone (alone) -> output heads
        # (concurrent); the backward of a site runs right after its forward, in the same phase.
        owner, phase_of = {}, {}
        for ph, key in ((0, "input_heads"), (2, "output_heads")):
            if key in model.model:
                for ti, head in enumerate(model.model[key]):
                    for m_ in head.modules():
                        owner[id(m_)] = ti
                        phase_of[id(m_)] = ph
        self.n_streams = (max(owner.values()) + 1) if (owner and concurrent_heads) else 1
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.n_streams)] if self.n_streams > 1 else []
        self.sites = []
        for mod, shp in sites:
            shp = (batch,) + shp[1:]
            self.sites.append((mod, rnd(*shp).requires_grad_(True), rnd(*shp), owner.get(id(mod), 0) % self.n_streams,
                               phase_of.get(id(mod), 1)))
        self.gdn_elems_per_image = sum(int(s_[1][0].numel()) for s_ in self.sites)
        zs, ys, ss = [(batch,) + shapes[k][1:] for k in ("z", "y", "s")]
        self.z = (torch.distributions.Laplace(0.0, 2.0).sample(zs).to(device)).requires_grad_(True)
        self.scales = torch.exp(torch.empty(ss, device=device).uniform_(-3.0, 4.16, generator=g)).requires_grad_(True)
        self.y = (rnd(*ys) * 3).requires_grad_(True)
        self.x = mm.synthetic_batch(TASKS, batch, device=device, seed=22)
        self.x_hat = {t: (v + 0.1 * rnd(*v.shape)).requires_grad_(True) for t, v in self.x.items()}
        self.eb, self.gc = c.entropy_bottleneck, c.gaussian_conditional
        # gradients of the rate-path parameters live in one flat bucket (exchanged when N > 1)
        self.eb_params = [p for n, p in self.eb.named_parameters() if n != "quantiles"]
        self.lv = list(model.loss_balancer.parameters())
        self.params = [p for s_ in self.sites for p in (s_[0].beta, s_[0].gamma)] + self.eb_params + self.lv
        self.bucket = mm.FlatGradBucket(self.params)
        self.loss_inputs = [self.z, self.y, self.scales] + list(self.x_hat.values()) + self.eb_params + self.lv

    def step(self, dist=None, world=1):
        torch, mm = self.torch, self.mm

        def site(mod, x, g):                  # K4: one GDN / IGDN site, forward + backward
            y = mod(x)
            _, gb, gg = torch.autograd.grad(y, [x, mod.beta, mod.gamma], g)
            mod.beta.grad.copy_(gb)
            mod.gamma.grad.copy_(gg)

        main = torch.cuda.current_stream()
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
        self.eb.train(), self.gc.train()
        mm.ops.noise_source.step()            # advance the device-side Philox stream (part of the captured graph)
        z_hat, z_lik = self.eb(self.z)        # K1 + K2
        y_hat, y_lik = self.gc(self.y, self.scales)  # K1 + K3
        lik = mm.compressors.LikelihoodDict(y=y_lik, z=z_lik)
        lik.log_sums = {"y": self.gc.last_log_likelihood_sums, "z": self.eb.last_log_likelihood_sums}
        loss, _ = self.model.rate_distortion_loss(self.x, self.x_hat, lik, "train")  # distortion + RD epilogue
        grads = torch.autograd.grad(loss, self.loss_inputs)
        for p, gr in zip(self.eb_params + self.lv, grads[-(len(self.eb_params) + len(self.lv)):]):
            p.grad.copy_(gr)
        if world > 1:                         # the path's one exchange step: rate-path parameter gradients
            dist.all_reduce(self.bucket.flat)
        return loss


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
    C ABI (effective beta / gamma precomputed) on alternating 839 MB inputs (>> 126 MB L2) and timed with CUDA
    events on the launching stream.  `traffic` is the ncu DRAM byte count of the same launch (profiles/)."""
    L = mm._lib.lib()
    prec = mm.ops.GDN_PRECISION[precision]
    big = sorted(harness.sites, key=lambda s: -s[1].numel())[:2]
    eff = []
    with torch.no_grad():
        for mod, x, g, _, _ in big:
            eff.append((mod.beta_reparam(mod.beta).clone(), mod.gamma_reparam(mod.gamma).clone(), x.detach(), g,
                        mod.inverse))
    x0 = eff[0][2]
    B, C = x0.shape[:2]
    HW = x0.numel() // (B * C)
    n = x0.numel()
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty_like(x0)
    dbeta, dgamma = torch.empty_like(eff[0][0]), torch.empty_like(eff[0][1])
    nbytes = int(L.mmnc_gdn_backward_workspace_bytes(B, C, HW, prec))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
    idx = [0]

    def fwd():
        b, gm, x, _, inv = eff[idx[0] % len(eff)]
        idx[0] += 1
        mm._lib.check(L.mmnc_gdn_forward(x.data_ptr(), B, C, HW, b.data_ptr(), gm.data_ptr(), int(inv), prec,
                                         y.data_ptr(), st))

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
    out = {"roofline": {
        "bound": "hbm", "kernel": "gdn backward (fused tcgen05, %s), %s; includes its ~10 us partial-reduce launch"
                                  % ({3: "TMA-fed pipelined", 2: "first generation"}.get(variant, "variant %d" % variant), shape),
        "achieved": 12.0 * n / t_b / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 12.0 * n / t_b / 1e9 / peak_gbs,
        "traffic": ncu_traffic[0], "peak_source": peak_src, "algorithmic_bytes_per_launch": 12 * n,
        "launch_ms": t_b * 1e3}}
    out["roofline_forward"] = {
        "bound": "hbm", "kernel": "gdn forward (tcgen05), " + shape,
        "achieved": 8.0 * n / t_f / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": 8.0 * n / t_f / 1e9 / peak_gbs,
        "traffic": ncu_traffic[1], "algorithmic_bytes_per_launch": 8 * n, "launch_ms": t_f * 1e3}
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
        Bq, Cq = shp[:2]
        HWq = x.numel() // (Bq * Cq)
        yq, dbq, dgq = torch.empty_like(x), torch.empty_like(b_), torch.empty_like(gm_)
        nbq = int(L.mmnc_gdn_backward_workspace_bytes(Bq, Cq, HWq, prec))
        wsq = torch.empty(nbq, dtype=torch.uint8, device=x.device)
        k = [0]

        def fq():
            k[0] += 1
            mm._lib.check(L.mmnc_gdn_forward(xs_[k[0] & 1].data_ptr(), Bq, Cq, HWq, b_.data_ptr(), gm_.data_ptr(),
                                             int(mod.inverse), prec, yq.data_ptr(), st))

        def bq():
            k[0] += 1
            mm._lib.check(L.mmnc_gdn_backward(xs_[k[0] & 1].data_ptr(), g.data_ptr(), Bq, Cq, HWq, b_.data_ptr(),
                                              gm_.data_ptr(), int(mod.inverse), prec, yq.data_ptr(), dbq.data_ptr(),
                                              dgq.data_ptr(), wsq.data_ptr(), nbq, st))

        tf_, tb_ = time_kernel(torch, fq, reps=10), time_kernel(torch, bq, reps=10)
        nq = x.numel()
        others.append({"layer": "%sGDN(%d) on %dx%d, batch %d" % ("I" if mod.inverse else "", Cq, shp[2], shp[3], Bq),
                       "forward_frac": 8.0 * nq / tf_ / 1e9 / peak_gbs, "forward_ms": tf_ * 1e3,
                       "backward_frac": 12.0 * nq / tb_ / 1e9 / peak_gbs, "backward_ms": tb_ * 1e3})
        del xs_, yq, wsq
    out["roofline_other_layers"] = others
    return out


# ----------------------------------------------------------------------------------------------------- CPU legs
def cpu_rate_path_images_per_s(cpu_batch: int, budget_s: float = 20.0):
    """The oracle (torch CPU ops in CompressAI's op order) on the same rate-path workload, bounded sample."""
    import torch

    from oracle import compressai_ref as R
    from oracle import reference_models as orm

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(21)
    ref = orm.ReferenceCompressor(3, TASKS, MODEL["latent_channels"], MODEL["conv_channels"], lmbda=MODEL["lmbda"])
    sites, shapes, hooks = [], {}, []
    for mod in ref.modules():
        if isinstance(mod, R.GDN):
            hooks.append(mod.register_forward_hook(lambda m, i, o: sites.append((m, tuple(i[0].shape)))))
    c = ref.model["compressor"]
    hooks.append(c.entropy_bottleneck.register_forward_hook(lambda m, i, o: shapes.__setitem__("z", tuple(i[0].shape))))
    hooks.append(c.gaussian_conditional.register_forward_hook(
        lambda m, i, o: shapes.update(y=tuple(i[0].shape), s=tuple(i[1].shape))))
    batch = orm.synthetic_batch(TASKS, cpu_batch)
    with torch.no_grad():
        ref(batch)
    for h in hooks:
        h.remove()
    data = [(m, torch.randn(*s).requires_grad_(True), torch.randn(*s)) for m, s in sites]
    z = (torch.randn(shapes["z"]) * 2).requires_grad_(True)
    y = (torch.randn(shapes["y"]) * 3).requires_grad_(True)
    sc = torch.exp(torch.empty(shapes["s"]).uniform_(-3, 4.16)).requires_grad_(True)
    x_hat = {t: (v + 0.1 * torch.randn_like(v)).requires_grad_(True) for t, v in batch.items()}
    ref.train()

    def step():
        for m, x, g in data:
            torch.autograd.grad(m(x), [x, m.beta, m.gamma], g)
        z_hat, z_lik = c.entropy_bottleneck(z)
        y_hat, y_lik = c.gaussian_conditional(y, sc)
        rec, _ = ref.multitask_reconstruction_loss(batch, x_hat)
        comp, _ = ref.multitask_compression_loss({"y": y_lik, "z": z_lik}, x_hat)
        (ref.lmbda * rec + comp).backward()

    step()
    t0, n = time.perf_counter(), 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 20:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": cpu_batch / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} timed passes of the same rate-path workload at batch {cpu_batch} (oracle, torch CPU ops, "
                      f"{torch.get_num_threads()} threads), {dt * 1e3:.0f} ms per pass"}


def run_reference(args):
    """--impl reference: the reference's training step restated on the host cores (oracle/reference_models.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import reference_models as orm

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(21)
    ref = orm.ReferenceCompressor(3, TASKS, MODEL["latent_channels"], MODEL["conv_channels"], lmbda=MODEL["lmbda"]).train()
    ref.configure_optimizers(total_steps=args.steps + args.warmup)
    B = args.cpu_batch
    batch = orm.synthetic_batch(TASKS, B)
    for _ in range(args.warmup):
        ref.training_step(batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss, _ = ref.training_step(batch)
        float(loss)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = B / dt
    sample = (f"full training step (heads + backbone convs + rate path + both Adam optimizers) at batch {B} on "
              f"{torch.get_num_threads()} host threads; oracle = CPU restatement of CompressAI 1.2.4 (parity unpinned)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s",
        "what": "the reference's whole training step on the host cores (conv heads + backbone + rate path + both Adam "
                "steps): the counterpart of the b200 arm's `e2e`; the rate path alone on the CPU is `cpu_baseline` of "
                "the b200 arm's line",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": B},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------- rANS leg
def rans_leg(torch, mm, model, device, n_images=1024):
    """compress/decompress throughput at the C2 latent shapes (y: 128, z: 300 symbols per image), strings
    returned to the host; CPU oracle timed on a bounded sample in both marshalling variants."""
    from oracle import compressai_ref as R

    c = model.model["compressor"]
    eb, gc = c.entropy_bottleneck, c.gaussian_conditional
    M, N = c.M, c.N
    g = torch.Generator(device=device).manual_seed(5)
    scales = torch.exp(torch.empty(n_images, M, 1, 1, device=device).uniform_(-3.0, 4.16, generator=g))
    y = torch.randn(n_images, M, 1, 1, device=device, generator=g) * scales
    y[torch.rand(y.shape, device=device, generator=g) < 0.01] *= 50.0
    z = torch.randn(n_images, N, 1, 1, device=device, generator=g) * 4

    def enc():
        idx = gc.build_indexes(scales)
        return gc.compress(y, idx), eb.compress(z), idx

    ys, zs, idx = enc()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ys, zs, idx = enc()
    torch.cuda.synchronize()
    t_enc = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3):
        yh = gc.decompress(ys, idx)
        zh = eb.decompress(zs, (1, 1))
    torch.cuda.synchronize()
    t_dec = (time.perf_counter() - t0) / 3
    ok = bool(torch.equal(yh, torch.round(y)))
    sym_bytes = n_images * (M + N) * 4
    out = {"images": n_images, "symbols_per_image": M + N, "roundtrip_ok": ok,
           "encode_MBps": sym_bytes / t_enc / 1e6, "decode_MBps": sym_bytes / t_dec / 1e6,
           "encode_Msym_s": n_images * (M + N) / t_enc / 1e6, "decode_Msym_s": n_images * (M + N) / t_dec / 1e6,
           "output_bytes": sum(map(len, ys)) + sum(map(len, zs))}
    # bpp match: actual bytes against the likelihood estimate on the same tensors
    gc.eval(), eb.eval()
    with torch.no_grad():
        est_bits = float(-torch.log2(gc(y, scales)[1]).sum() - torch.log2(eb(z)[1]).sum())
    out["bits_actual_over_estimated"] = 8 * out["output_bytes"] / est_bits
    # CPU oracle on 8 images, CompressAI-faithful marshalling and lean marshalling
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
    return out


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
    torch.manual_seed(21)
    B = args.batch

    model = mm.build_compressor(MODEL["model_type"], TASKS, MODEL["latent_channels"], MODEL["conv_channels"],
                                lmbda=MODEL["lmbda"])
    for m in model.modules():
        if isinstance(m, mm.GDN):
            m.precision = args.precision
    model.update_bottleneck_values()  # CPU, before .to(device): the reference's order (src/compress.py:101-105)
    model.to(device)
    model.train()
    dp = mm.DataParallel(model) if world > 1 else None
    model.configure_optimizers(total_steps=10 * (args.steps + args.warmup))

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
    harness = RatePathHarness(mm, model, B, device, torch, concurrent_heads=not args.serial_heads)
    mm.ops.noise_source.enable_device_state(device, seed=21)  # Philox (seed, offset) on the device: graph-safe
    run_step = lambda: harness.step(dist, world)  # noqa: E731
    graph = None
    if not args.no_graph:
        # The step is ~600 small launches: without a graph the host cannot issue them as fast as the GPU retires
        # them.  Capture once, replay per step (fresh noise each replay through the device-side Philox state).
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
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = mm.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    t_rate = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / args.steps
    launches = (launches_per_step * args.steps) if graph is not None else (mm.launch_count() - launches0)
    mm.ops.noise_source.disable_device_state()
    del graph
    value = world * B / t_rate

    # -------- e2e: public API, host batch, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        del harness.bucket
        host = mm.synthetic_batch(TASKS, B, seed=21 + rank, pin_memory=True)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        if dp is not None:  # re-home the gradients of the whole model into DataParallel's bucket
            dp.bucket = mm.FlatGradBucket(list(model.get_main_parameters()) + list(model.loss_balancer.parameters()))

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
            loss = model.training_step(dev_batch)
            pending[0] = upload()
            return float(loss.item())  # D2H read of the step's result

        for _ in range(max(args.warmup, 3)):
            e2e_step()
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            last = e2e_step()
        e1.record()
        barrier()
        t_e2e = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)) / args.steps
        e2e = {"value": world * B / t_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": t_e2e * 1e3, "last_loss": last,
               "what": "compressor.training_step(batch): H2D of one pinned batch per step (copy stream, prefetched one step "
                       "ahead), cuDNN convs (TF32 allowed, torch default) + mmnc kernels, backward, gradient all-reduce "
                       "(N>1), both Adam steps, loss.item()"}
    clocks = sampler.stop() if rank == 0 else None

    # -------- roofline of the dominant kernel, CPU baseline, rANS leg (rank 0 only; outside the timed regions)
    extra = {}
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        harness2 = harness
        extra.update(roofline_of_gdn(torch, mm, harness2, peak, src, args.precision))
        if world == 1 and not args.no_rans:
            try:
                extra["rans"] = rans_leg(torch, mm, model, device)
            except Exception as ex:  # the leg is auxiliary: never lose the main line
                extra["rans"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu:
            extra["cpu_baseline"] = cpu_rate_path_images_per_s(args.cpu_batch)
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_rate * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "gdn_precision": args.precision,
                       "gdn_elements_per_image": harness.gdn_elems_per_image,
                       "launch": "eager" if args.no_graph else "CUDA graph replay of one captured step",
                       "head_streams": harness.n_streams,
                       "l2_policy": "inputs larger than L2 (4.7 GB of GDN activations per step vs 126 MB L2)"},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
        }
        # whole-step view: algorithmic GDN traffic of one step (8 B forward + 12 B backward per element; the entropy
        # models and distortion terms add < 1 %) over the measured step time, against the same measured peak
        if "roofline" in extra:
            step_bytes = 20.0 * harness.gdn_elems_per_image * B
            pk = extra["roofline"]["peak"]
            line["step_roofline"] = {"bound": "hbm", "achieved": step_bytes / t_rate / 1e9, "peak": pk, "unit": "GB/s",
                                     "frac": step_bytes / t_rate / 1e9 / pk, "algorithmic_bytes_per_step": step_bytes}
        line.update(extra)
        if "cpu_baseline" not in line:
            line["cpu_baseline"] = None
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
