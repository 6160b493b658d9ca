"""Reproducer for the MN-major B operand experiment of gdn_tc_bwd2.cu (MMNC_BWD2_MN=1): decomposes the kernel's dx into
f + 2 x t and reports which hypothesis for t explains it (true t, t built from gamma without the transpose, t = 0).
usage: MMNC_GDN_BWD=v2 MMNC_BWD2_MN=1..4 python tools/mn_major_probe.py     (without MMNC_BWD2_MN: the shipped path)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mmnc_b200 as mm
dev = "cuda:0"
torch.manual_seed(0)
B, C, H, W = 10, 50, 64, 64
x = torch.randn(B, C, H, W, device=dev); g = torch.randn(B, C, H, W, device=dev)
beta = 1 + torch.rand(C, device=dev) * 0.5
gamma = 0.1 * torch.eye(C, device=dev) + torch.rand(C, C, device=dev) * 0.05
xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
y = mm.ops.gdn(xr, br, gr, False, "tf32")
dx, db, dg = torch.autograd.grad(y, [xr, br, gr], g)
x64, b64, g64 = x.double().requires_grad_(True), beta.double(), gamma.double()
n = torch.nn.functional.conv2d(x64 * x64, g64.reshape(C, C, 1, 1), b64)
y64 = x64 * torch.rsqrt(n)
(wdx,) = torch.autograd.grad(y64, [x64], g.double())
# decomposition: dx = f + 2 x t ; f = g n^-1/2 ; t = u gamma with u = -1/2 g x n^-3/2
f = g.double() * n.detach() ** -0.5
u = -0.5 * g.double() * x.double() * n.detach() ** -1.5
t_true = torch.einsum("bihw,ik->bkhw", u, g64)
t_T = torch.einsum("bihw,ki->bkhw", u, g64)      # what using gamma instead of gamma^T would give
xd = x.double()
def res(t):  # how well dx = f + 2 x t explains the kernel's dx
    return ((dx.double() - (f + 2 * xd * t)).norm() / (2 * xd * t_true).norm()).item()
# other hypotheses: t built from gamma (not transposed); t = 0; t from gamma with rows/cols of 4x8 core blocks swapped
print("residual with t_true       :", res(t_true))
print("residual with gamma (no T) :", res(t_T))
print("residual with t = 0        :", res(torch.zeros_like(t_true)))
