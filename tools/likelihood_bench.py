"""EB / GC / distortion kernels at the declared roofline shape and at the reference's shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
dev = torch.device("cuda:0"); peak = 6547.5
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
for name, zs, ys, ss in (("roofline shape", (1024, 512, 4, 4), (1024, 192, 16, 16), (1024, 192, 16, 16)),
                         ("reference C2 B=64", (64, 300, 1, 1), (64, 128, 1, 1), (64, 128, 4, 4)),
                         ("reference C4 B=256", (256, 512, 1, 1), (256, 190, 1, 1), (256, 190, 4, 4))):
    eb = mm.EntropyBottleneck(zs[1]).to(dev).train(); gc = mm.GaussianConditional(None).to(dev).train()
    z = (torch.randn(*zs, device=dev) * 3).requires_grad_(True)
    y = (torch.randn(*ys, device=dev) * 3).requires_grad_(True)
    sc = torch.exp(torch.empty(*ss, device=dev).uniform_(-3, 4.16)).requires_grad_(True)
    packed = eb.packed_parameters().detach().requires_grad_(True); med = eb._get_medians().detach().reshape(-1)
    with torch.no_grad():
        tf = t(lambda: mm.ops.entropy_bottleneck_forward(z, packed, med, True, 1e-9, seed=1))
        tg = t(lambda: mm.ops.gaussian_conditional_forward(y, sc, None, True, 0.11, 1e-9, seed=1))
    def ebfb():
        o, l, s = mm.ops.entropy_bottleneck_forward(z, packed, med, True, 1e-9, seed=1)
        torch.autograd.grad(s.sum(), [z, packed])
    def gcfb():
        o, l, s = mm.ops.gaussian_conditional_forward(y, sc, None, True, 0.11, 1e-9, seed=1)
        torch.autograd.grad(s.sum(), [y, sc])
    tfb, tgb = t(ebfb) - tf, t(gcfb) - tg
    nz, nl, r = z.numel(), sc.numel(), sc.numel() // y.numel()
    print(f"{name:20s} EB fwd {tf*1e6:8.1f} us {12*nz/tf/1e9:6.0f} GB/s ({12*nz/tf/1e9/peak:.3f}) bwd {tfb*1e6:8.1f} us {16*nz/tfb/1e9:6.0f} GB/s | "
          f"GC fwd {tg*1e6:8.1f} us {(8+8/r)*nl/tg/1e9:6.0f} GB/s ({(8+8/r)*nl/tg/1e9/peak:.3f}) bwd {tgb*1e6:8.1f} us {(12+12/r)*nl/tgb/1e9:6.0f} GB/s")
a = torch.rand(64, 3, 256, 256, device=dev); b = (a + 0.1 * torch.randn_like(a)).requires_grad_(True)
with torch.no_grad(): td = t(lambda: mm.ops.distortion(b, a, "mse"))
tdb = t(lambda: torch.autograd.grad(mm.ops.distortion(b, a, "mse"), [b])) - td
print(f"distortion (64,3,256,256): fwd {td*1e6:.1f} us {8*a.numel()/td/1e9:.0f} GB/s ({8*a.numel()/td/1e9/peak:.3f})  bwd {tdb*1e6:.1f} us {12*a.numel()/tdb/1e9:.0f} GB/s")
