"""One launch each of the EB / GC kernels at the declared roofline shape, for `ncu --set full` (profiles/)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
dev = torch.device("cuda:0")
eb = mm.EntropyBottleneck(512).to(dev).train(); gc = mm.GaussianConditional(None).to(dev).train()
z = (torch.randn(1024, 512, 4, 4, device=dev) * 3).requires_grad_(True)
y = (torch.randn(1024, 192, 16, 16, device=dev) * 3).requires_grad_(True)
sc = torch.exp(torch.empty(1024, 192, 16, 16, device=dev).uniform_(-3, 4.16)).requires_grad_(True)
packed = eb.packed_parameters().detach().requires_grad_(True); med = eb._get_medians().detach().reshape(-1)
for _ in range(2):
    o, l, s = mm.ops.entropy_bottleneck_forward(z, packed, med, True, 1e-9, seed=1)
    torch.autograd.grad(s.sum(), [z, packed])
    o, l, s = mm.ops.gaussian_conditional_forward(y, sc, None, True, 0.11, 1e-9, seed=1)
    torch.autograd.grad(s.sum(), [y, sc])
torch.cuda.synchronize()
print("ok")
