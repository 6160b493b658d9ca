import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
dev = "cuda:0"
torch.manual_seed(0)
def ref(x, g, beta, gamma, inverse):
    x = x.double().requires_grad_(True); beta = beta.double().requires_grad_(True); gamma = gamma.double().requires_grad_(True)
    C = x.shape[1]
    n = torch.nn.functional.conv2d(x * x, gamma.reshape(C, C, 1, 1), beta)
    y = x * (torch.sqrt(n) if inverse else torch.rsqrt(n))
    return torch.autograd.grad(y, [x, beta, gamma], g.double())
for C, H in ((50, 64),):
    for kind in ("random", "symmetric", "diagonal"):
        x = torch.randn(2, C, H, H, device=dev); g = torch.randn(2, C, H, H, device=dev)
        beta = 1 + torch.rand(C, device=dev) * 0.5
        gamma = 0.1 * torch.eye(C, device=dev)
        if kind == "random": gamma = gamma + torch.rand(C, C, device=dev) * 0.05
        if kind == "symmetric":
            r = torch.rand(C, C, device=dev) * 0.05; gamma = gamma + (r + r.t()) / 2
        for inv in (False, True):
            xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
            out = {}
            for prec in ("fp32", "tf32"):
                out[prec] = torch.autograd.grad(mm.ops.gdn(xr, br, gr, inv, prec), [xr, br, gr], g)
            want = ref(x, g, beta, gamma, inv)
            errs = []
            for prec in ("fp32", "tf32"):
                errs.append(" ".join("%.1e" % ((a.double() - w).abs().max() / w.abs().max()).item() for a, w in zip(out[prec], want)))
            print(f"C={C} {kind:9s} inv={inv}: fp32 [dx dbeta dgamma] {errs[0]} | tf32 {errs[1]}")
