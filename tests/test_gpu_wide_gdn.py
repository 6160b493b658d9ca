"""GDN / IGDN with 129 .. 256 channels (IGDN(256) of BASELINE config C3: mtc.py:146-172 with c = 128 doubled by the
mixed-latent heads) on the tensor cores: gdn_tc_wide.cu against the float64 oracle at the stated single-pass TF32 bars
(forward rtol 1e-3; backward 2e-3 of the largest entry, like the <= 128 channel kernels)."""
import pytest
import torch

import mmnc_b200 as mm
from oracle import compressai_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(C, inverse, seed=0):
    g = torch.Generator().manual_seed(seed)
    ref = R.GDN(C, inverse=inverse)
    with torch.no_grad():
        ref.gamma.add_(torch.rand(C, C, generator=g) * 0.05)
        ref.beta.add_(torch.rand(C, generator=g) * 0.5)
    ours = mm.GDN(C, inverse=inverse, precision="tf32")
    ours.load_state_dict(ref.state_dict())
    return ours.to(DEV), ref


def _variants(x):
    B, C = x.shape[:2]
    HW = x.numel() // (B * C)
    L, p = mm._lib.lib(), mm.ops.GDN_PRECISION["tf32"]
    return (int(L.mmnc_gdn_forward_variant(x.data_ptr(), x.data_ptr(), B, C, HW, p)),
            int(L.mmnc_gdn_backward_variant(x.data_ptr(), x.data_ptr(), B, C, HW, p)))


# 4096 .. 40960 pixels = 32 .. 320 tiles on 148 CTAs (CTAs with 1, 2 and 3 tiles); 65 x 64: a ragged last tile and
# pixels of two images in one tile; C = 129 / 160 / 192 / 200: padded channels and fewer K chunks than 8
FWD_SHAPES = [(1, 256, 64, 64), (4, 256, 32, 32), (10, 256, 64, 64), (2, 192, 64, 64), (3, 160, 48, 40), (1, 129, 65, 64),
              (5, 200, 32, 33), (2, 255, 64, 32)]


@pytest.mark.parametrize("shape", FWD_SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
def test_wide_gdn_forward(shape, inverse):
    torch.manual_seed(23)
    ours, ref = _pair(shape[1], inverse)
    x = torch.randn(*shape)
    xd = x.to(DEV)
    assert _variants(xd)[0] == 6, "shape should take the wide-layer kernel"
    before = mm.launch_count()
    y = ours(xd)
    torch.cuda.synchronize()
    assert mm.launch_count() - before == 2  # pack gamma, streamed-operand forward
    want = ref(x)
    assert torch.allclose(y.detach().cpu(), want, rtol=1e-3, atol=1e-4), (y.detach().cpu() - want).abs().max()
    assert torch.equal(ours(xd), y), "deterministic"
    assert torch.equal(xd.cpu(), x)


BWD_SHAPES = [(1, 256, 64, 64), (4, 256, 32, 32), (10, 256, 64, 64), (2, 192, 64, 64), (3, 160, 32, 64), (1, 129, 64, 64),
              (5, 200, 32, 32), (300, 256, 4, 8)]


@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
def test_wide_gdn_backward(shape, inverse):
    torch.manual_seed(29)
    C = shape[1]
    ours, ref = _pair(C, inverse)
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    xd, gd = x.to(DEV).requires_grad_(True), g.to(DEV)
    assert _variants(xd) == (6, 6)
    before = mm.launch_count()
    ours(xd).backward(gd)
    torch.cuda.synchronize()
    assert mm.launch_count() - before == 2 + 4  # forward pair; pack, dx, d gamma, partial reduce
    x64 = x.double().requires_grad_(True)
    refd(x64).backward(g.double())

    def close(a, b, tol):
        return ((a.cpu().double() - b).abs().max() / b.abs().max()).item() <= tol

    assert close(xd.grad, x64.grad, 2e-3), ((xd.grad.cpu().double() - x64.grad).abs().max(), x64.grad.abs().max())
    assert close(ours.beta.grad, refd.beta.grad, 2e-3)
    assert close(ours.gamma.grad, refd.gamma.grad, 2e-3)
    first = (xd.grad.clone(), ours.beta.grad.clone(), ours.gamma.grad.clone())
    xd.grad, ours.beta.grad, ours.gamma.grad = None, None, None
    ours(xd).backward(gd)
    assert all(torch.equal(a, b) for a, b in zip(first, (xd.grad, ours.beta.grad, ours.gamma.grad))), "deterministic"


def test_wide_gdn_falls_back_when_shape_does_not_suit():
    """Small problems, pixel counts that are not a multiple of 32 (backward: TMA boxes) and fp32 requests run the fp32
    SIMT kernels, which the float64 oracle holds to 2e-5."""
    torch.manual_seed(31)
    ours, ref = _pair(256, False)
    x = torch.randn(2, 256, 8, 8)  # 128 pixels: below the wide kernels' minimum
    assert _variants(x.to(DEV)) == (1, 1)
    assert torch.allclose(ours(x.to(DEV)).cpu(), ref(x), rtol=2e-5, atol=1e-6)
    x = torch.randn(2, 256, 50, 50).to(DEV)  # 2500 pixels per image: forward wide, backward SIMT
    assert _variants(x) == (6, 1)
    xg = x.clone().requires_grad_(True)
    ours(xg).sum().backward()
    assert torch.isfinite(xg.grad).all()


@pytest.mark.parametrize("shape,inverse", [((32, 256, 64, 64), True), ((64, 256, 32, 32), True), ((16, 256, 64, 64), False)])
def test_wide_gdn_full_size_properties(shape, inverse):
    """IGDN(256) @ 64 x 64 and @ 32 x 32 as BASELINE config C3 runs them (half / all of its batch of 64): size-independent
    properties - odd symmetry and batch independence hold EXACTLY (the same roundings in the same order), the parameter
    gradients of the two halves of the batch add up to the whole - and torch's own fp32 ops on the same device."""
    torch.manual_seed(37)
    B, C, H, W = shape
    ours, _ = _pair(C, inverse)
    beta, gamma = ours.beta_reparam(ours.beta).detach(), ours.gamma_reparam(ours.gamma).detach()
    x = torch.randn(*shape, device=DEV)
    g = torch.randn(*shape, device=DEV)
    assert _variants(x) == (6, 6)

    def run(xi, gi):
        xr, br, gr = xi.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
        y = mm.ops.gdn(xr, br, gr, inverse, "tf32")
        return (y.detach(),) + torch.autograd.grad(y, [xr, br, gr], gi)

    y, dx, db, dg = run(x, g)
    y2, dx2, db2, dg2 = run(-x, g)
    assert torch.equal(y2, -y) and torch.equal(dx2, dx) and torch.equal(db2, -db) and torch.equal(dg2, -dg)
    y3, dx3, _, _ = run(x[3:7].contiguous(), g[3:7].contiguous())
    assert torch.equal(y3, y[3:7]) and torch.equal(dx3, dx[3:7])
    _, _, dba, dga = run(x[: B // 2].contiguous(), g[: B // 2].contiguous())
    _, _, dbb, dgb = run(x[B // 2:].contiguous(), g[B // 2:].contiguous())
    assert ((dga + dgb - dg).abs().max() / dg.abs().max()).item() < 1e-5
    assert ((dba + dbb - db).abs().max() / db.abs().max()).item() < 1e-5
    xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        norm = torch.nn.functional.conv2d(xr * xr, gr.reshape(C, C, 1, 1), br)
        want = xr * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
        wdx, wdb, wdg = torch.autograd.grad(want, [xr, br, gr], g)
    assert torch.allclose(y, want.detach(), rtol=1e-3, atol=1e-4)
    for got, ref in ((dx, wdx), (db, wdb), (dg, wdg)):
        assert ((got - ref).abs().max() / ref.abs().max()).item() <= 2e-3


def test_wide_gdn_cluster_multicast_path_matches_default():
    """MMNC_GDN_WIDE_CLUSTER=2 (2-CTA clusters, gamma chunks multicast into both CTAs' rings; not the default - it measured
    no faster) is read once per process, so it runs in a child process: same results as the default path, bit for bit."""
    import os
    import subprocess
    import sys

    code = (
        "import torch, mmnc_b200 as mm\n"
        "torch.manual_seed(41)\n"
        "m = mm.GDN(256, inverse=True, precision='tf32').to('cuda:0')\n"
        "with torch.no_grad():\n"
        "    m.gamma.add_(torch.rand(256, 256, device='cuda:0') * 0.05)\n"
        "x = torch.randn(5, 256, 64, 32, device='cuda:0', requires_grad=True)\n"
        "g = torch.randn(5, 256, 64, 32, device='cuda:0')\n"
        "y = m(x); y.backward(g)\n"
        "torch.save([t.cpu() for t in (y.detach(), x.grad, m.beta.grad, m.gamma.grad)], __import__('sys').argv[1])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for cs in ("1", "2"):
        path = os.path.join("/tmp", f"mmnc_wide_cluster_{cs}_{os.getpid()}.pt")
        env = dict(os.environ, MMNC_GDN_WIDE_CLUSTER=cs, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, cwd=root, timeout=300)
        outs.append(torch.load(path))
        os.remove(path)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
