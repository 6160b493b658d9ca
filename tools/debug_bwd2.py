import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
dev = "cuda:0"
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
C, H = 16, 64
i0 = int(os.environ.get("I0", "3"))
x = torch.ones(1, C, H, H, device=dev)
g = torch.zeros(1, C, H, H, device=dev); g[:, i0] = 1.0
beta = torch.ones(C, device=dev)
ii, kk = torch.meshgrid(torch.arange(C), torch.arange(C), indexing="ij")
gamma = ((ii * 100 + kk).float() * 1e-4).to(dev)
xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
dx = torch.autograd.grad(mm.ops.gdn(xr, br, gr, False, "tf32"), [xr, br, gr], g)[0]
n = beta + gamma.sum(1)                      # x^2 = 1
rs = n.rsqrt()
u = -0.5 * 1.0 * 1.0 * rs[i0] ** 3           # only channel i0
f = torch.zeros(C, device=dev); f[i0] = rs[i0]
for pix in (0, 1, 37, 129):
    t = (dx[0, :, pix // H, pix % H] - f) / 2.0
    print("pixel", pix, "gamma read (x1e4):", (t / u * 1e4).round().int().tolist())
print("expected row i0:", (gamma[i0] * 1e4).round().int().tolist())
