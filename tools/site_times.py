"""Per-site timing of the rate-path harness (which GDN sites / entropy-model calls take the time)."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
import bench

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = mm.build_compressor(3, bench.TASKS, 128, 100, lmbda=1e-2).to(dev).train()
h = bench.RatePathHarness(mm, model, B, dev, torch)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rows = collections.OrderedDict()
for mod, x, g in h.sites:
    key = (tuple(x.shape[1:]), mod.inverse)
    with torch.no_grad():
        tf = t(lambda: mod(x))
    tfb = t(lambda: torch.autograd.grad(mod(x), [x, mod.beta, mod.gamma], g))
    r = rows.setdefault(key, [0, 0.0, 0.0]); r[0] += 1; r[1] += tf; r[2] += tfb - tf
tot_f = tot_b = 0
for (shape, inv), (n, tf, tb) in rows.items():
    el = shape[0] * shape[1] * shape[2] * B * n
    print(f"{'IGDN' if inv else 'GDN '} {str(shape):18s} x{n}: fwd {tf:7.3f} ms ({8*el/tf/1e6:6.0f} GB/s)  bwd {tb:7.3f} ms ({12*el/tb/1e6:6.0f} GB/s)")
    tot_f += tf; tot_b += tb
print(f"GDN total fwd {tot_f:.2f} ms, bwd {tot_b:.2f} ms")
def rest():
    h.eb.train(); h.gc.train()
    z_hat, z_lik = h.eb(h.z); y_hat, y_lik = h.gc(h.y, h.scales)
    lik = mm.compressors.LikelihoodDict(y=y_lik, z=z_lik)
    lik.log_sums = {"y": h.gc.last_log_likelihood_sums, "z": h.eb.last_log_likelihood_sums}
    loss, _ = h.model.rate_distortion_loss(h.x, h.x_hat, lik, "train")
    torch.autograd.grad(loss, h.loss_inputs)
print(f"EB + GC + distortion + RD epilogue fwd+bwd: {t(rest):.3f} ms")
print(f"whole step eager: {t(lambda: h.step()):.2f} ms")
