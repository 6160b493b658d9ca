"""Parity tests proper: the CUDA path, called through the C ABI (via the package's ops), against the oracle on
identical seeded inputs.  Bit-exact for symbols / indexes / tables / byte strings; tiered 1e-5 relative for
likelihoods (helpers.assert_likelihood_close); stated tolerances for GDN.  Run with `-m gpu` on a B200."""
import json
import os

import numpy as np
import pytest
import torch

import mmnc_b200 as mm
from helpers import assert_likelihood_close, eb_double, load_golden, perturb_eb_
from oracle import compressai_ref as R
from oracle import native
from oracle import reference_models as orm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair_eb(C, perturbed=True, form="sign", seed=7):
    ref = R.EntropyBottleneck(C, likelihood_form=form)
    if perturbed:
        perturb_eb_(ref, seed)
    ours = mm.EntropyBottleneck(C, likelihood_form=form)
    ours.load_state_dict(ref.state_dict())
    return ours, ref


@pytest.fixture(scope="module")
def gc_pair():
    ours, ref = mm.GaussianConditional(None), R.GaussianConditional(None)
    ours.update_scale_table(mm.get_scale_table())
    ref.update_scale_table(R.get_scale_table())
    return ours.to(DEV), ref


# ------------------------------------------------------------------------------------------------ a2
def test_quantize_family_bit_exact():
    torch.manual_seed(0)
    x = torch.randn(3, 5, 4, 6) * 9
    x[0, 0, 0, :6] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5])  # round-half-to-even cases
    med = torch.randn(5, 1, 1)
    em, er = mm.EntropyModel(), R.EntropyModel()
    xd, md = x.to(DEV), med.to(DEV)
    assert torch.equal(em.quantize(xd, "dequantize", md.reshape(1, 5, 1, 1)).cpu(), er.quantize(x, "dequantize", med))
    assert torch.equal(em.quantize(xd, "symbols", md.reshape(1, 5, 1, 1)).cpu(), er.quantize(x, "symbols", med))
    assert torch.equal(em.quantize(xd, "dequantize").cpu(), torch.round(x))
    full = torch.randn_like(x)
    assert torch.equal(em.quantize(xd, "symbols", full.to(DEV)).cpu(), er.quantize(x, "symbols", full))
    sym = er.quantize(x, "symbols", med)
    assert torch.equal(em.dequantize(sym.to(DEV), md.reshape(1, 5, 1, 1)).cpu(), er.dequantize(sym, med.expand(3, 5, 1, 1)))
    n = em.quantize(xd, "noise").cpu() - x
    assert n.min() >= -0.5001 and n.max() <= 0.5001 and abs(n.mean()) < 0.05 and abs(n.var() - 1 / 12) < 0.02
    noise = torch.rand_like(x) - 0.5
    assert torch.equal(mm.ops.quantize_noise(xd, noise=noise.to(DEV)).cpu(), x + noise)
    with pytest.raises(ValueError):
        em.quantize(xd, "nearest")


# ------------------------------------------------------------------------------------------------ a3
@pytest.mark.parametrize("shape", [(64, 30, 1, 1), (5, 6, 3, 2), (2, 33, 4, 4), (7, 1, 1, 1), (3, 9)])
@pytest.mark.parametrize("perturbed", [False, True])
def test_eb_forward_eval_and_train(shape, perturbed):
    torch.manual_seed(1)
    C = shape[1]
    ours, ref = _pair_eb(C, perturbed)
    ours.to(DEV)
    z = torch.randn(*shape) * 5
    z.view(-1)[:2] = torch.tensor([70.0, -70.0])
    refd = eb_double(ref)
    # eval: dequantize with medians
    ours.eval(), ref.eval(), refd.eval()
    out, lik = ours(z.to(DEV))
    out_r, lik_r = ref(z)
    _, lik_d = refd(z.double())
    assert torch.equal(out.cpu(), out_r)
    assert_likelihood_close(lik, lik_r, want64=lik_d, what="EB eval")
    assert torch.allclose(ours.last_log_likelihood_sums.cpu().double(),
                          torch.log(lik_r.double()).transpose(0, 1).reshape(C, -1).sum(1), rtol=2e-5, atol=1e-4)
    # train with injected noise
    ours.train(), ref.train(), refd.train()
    noise = torch.rand(*shape) - 0.5
    out, lik = ours(z.to(DEV), noise=noise.to(DEV))
    out_r, lik_r = ref(z, noise=noise)
    _, lik_d = refd(z.double(), noise=noise.double())
    assert torch.equal(out.cpu(), out_r)
    assert_likelihood_close(lik, lik_r, want64=lik_d, what="EB train")


def test_eb_golden_fixture():
    fx = load_golden()
    ours = mm.EntropyBottleneck(6)
    ours.load_state_dict({k: torch.from_numpy(v) for k, v in fx["eb_state"].item().items()})
    ours.to(DEV).eval()
    z = torch.from_numpy(fx["eb_z"]).to(DEV)
    out, lik = ours(z)
    assert np.array_equal(out.detach().cpu().numpy(), fx["eb_eval_out"])
    assert_likelihood_close(lik, torch.from_numpy(fx["eb_eval_lik"]), rtol=2e-5, what="EB golden eval")
    assert [s for s in ours.compress(z)] == [w.tobytes() for w in fx["eb_strings"]]
    ours.train()
    out, lik = ours(z, noise=torch.from_numpy(fx["eb_noise"]).to(DEV))
    assert np.array_equal(out.detach().cpu().numpy(), fx["eb_train_out"])
    assert_likelihood_close(lik, torch.from_numpy(fx["eb_train_lik"]), rtol=2e-5, what="EB golden train")
    assert abs(ours.loss().item() - float(fx["eb_aux_loss"])) <= 1e-5 * abs(float(fx["eb_aux_loss"]))


def test_eb_plain_form_and_philox_noise():
    torch.manual_seed(2)
    ours, ref = _pair_eb(8, True, form="plain")
    ours.to(DEV).eval(), ref.eval()
    z = torch.randn(16, 8, 2, 2) * 3
    _, lik = ours(z.to(DEV))
    _, lik_r = ref(z)
    assert_likelihood_close(lik, lik_r, rtol=1e-4, tail_rtol=5e-2, floor_atol=1e-8, what="EB plain")
    ours.train()
    torch.manual_seed(5)
    a, _ = ours(z.to(DEV))
    torch.manual_seed(5)
    b, _ = ours(z.to(DEV))
    assert torch.equal(a, b), "same torch seed -> same Philox noise"
    d = (a.cpu() - z)
    assert d.abs().max() <= 0.5 and d.std() > 0.2


def test_eb_backward_matches_autograd():
    torch.manual_seed(3)
    C = 6
    ours, ref = _pair_eb(C, True)
    refd = eb_double(ref).train()
    ours.to(DEV).train()
    z = (torch.randn(9, C, 2, 3) * 4)
    noise = torch.rand_like(z) - 0.5
    g_lik, g_out = torch.randn_like(z), torch.randn_like(z)
    zd = z.to(DEV).requires_grad_(True)
    out, lik = ours(zd, noise=noise.to(DEV))
    lnsum = ours.last_log_likelihood_sums
    w = torch.randn(C)
    ((lik * g_lik.to(DEV)).sum() + (out * g_out.to(DEV)).sum() + (lnsum * w.to(DEV)).sum()).backward()
    z64 = z.double().requires_grad_(True)
    out_r, lik_r = refd(z64, noise=noise.double())
    ln_r = torch.log(lik_r).transpose(0, 1).reshape(C, -1).sum(1)
    ((lik_r * g_lik.double()).sum() + (out_r * g_out.double()).sum() + (ln_r * w.double()).sum()).backward()
    assert torch.allclose(zd.grad.cpu().double(), z64.grad, rtol=2e-3, atol=1e-5)
    for name, p in ours.named_parameters():
        if name == "quantiles":
            continue
        want = dict(refd.named_parameters())[name].grad
        assert torch.allclose(p.grad.cpu().double(), want, rtol=2e-3, atol=2e-4), name


def test_eb_aux_loss_and_gradient():
    ours, ref = _pair_eb(12, True)
    ours.to(DEV)
    l, lr = ours.loss(), ref.loss()
    assert abs(l.item() - lr.item()) <= 1e-5 * abs(lr.item())
    l.backward(), lr.backward()
    assert torch.allclose(ours.quantiles.grad.cpu(), ref.quantiles.grad, rtol=1e-4, atol=1e-6)
    assert all(p.grad is None for n, p in ours.named_parameters() if n != "quantiles")


def test_eb_empty_batch():
    ours, _ = _pair_eb(4, False)
    ours.to(DEV).eval()
    out, lik = ours(torch.zeros(0, 4, 2, 2, device=DEV))
    assert out.shape == (0, 4, 2, 2) and lik.shape == (0, 4, 2, 2)


# ------------------------------------------------------------------------------------------------ a5
@pytest.mark.parametrize("yshape,sshape", [((3, 8, 1, 1), (3, 8, 4, 4)), ((4, 6, 4, 4), (4, 6, 4, 4)),
                                           ((2, 5, 3, 7), (2, 5, 3, 7)), ((2, 4, 1, 1), (2, 4, 3, 5)),
                                           ((2, 4, 1, 3), (2, 4, 2, 3))])
def test_gc_forward_shapes(gc_pair, yshape, sshape):
    ours, ref = gc_pair
    torch.manual_seed(4)
    scales = torch.exp(torch.empty(*sshape).uniform_(np.log(0.05), np.log(64)))
    y = torch.randn(*yshape) * 3
    y.view(-1)[0] = 800.0
    ours.eval(), ref.eval()
    yh, yl = ours(y.to(DEV), scales.to(DEV))
    yh_r, yl_r = ref(y, scales)
    refd = R.GaussianConditional(None).double().eval()
    _, yl_d = refd(y.double(), scales.double())
    assert yh.shape == yh_r.shape and yl.shape == yl_r.shape
    assert torch.equal(yh.cpu(), yh_r)
    assert_likelihood_close(yl, yl_r, want64=yl_d, what="GC eval")
    C = sshape[1]
    assert torch.allclose(ours.last_log_likelihood_sums.cpu().double(),
                          torch.log(yl_r.double()).transpose(0, 1).reshape(C, -1).sum(1), rtol=2e-5, atol=1e-4)
    ours.train(), ref.train()
    noise = torch.rand(*yshape) - 0.5
    yh, yl = ours(y.to(DEV), scales.to(DEV), noise=noise.to(DEV))
    yh_r, yl_r = ref(y, scales, noise=noise)
    assert torch.equal(yh.cpu(), yh_r)
    assert_likelihood_close(yl, yl_r, what="GC train", rtol=2e-5)


def test_gc_golden_and_means(gc_pair):
    ours, ref = gc_pair
    fx = load_golden()
    ours.eval(), ref.eval()
    yh, yl = ours(torch.from_numpy(fx["gc_y_bcast"]).to(DEV), torch.from_numpy(fx["gc_scales"]).to(DEV))
    assert np.array_equal(yh.cpu().numpy(), fx["gc_yhat_bcast"])
    assert_likelihood_close(yl, torch.from_numpy(fx["gc_lik_bcast"]), rtol=2e-5, what="GC golden")
    torch.manual_seed(6)
    y, sc, mu = torch.randn(2, 3, 4, 4) * 4, torch.rand(2, 3, 4, 4) * 3 + 0.05, torch.randn(2, 3, 4, 4)
    yh, yl = ours(y.to(DEV), sc.to(DEV), means=mu.to(DEV))
    yh_r, yl_r = ref(y, sc, means=mu)
    assert torch.equal(yh.cpu(), yh_r)
    assert_likelihood_close(yl, yl_r, rtol=2e-5, what="GC means")


@pytest.mark.parametrize("yshape,sshape", [((3, 8, 1, 1), (3, 8, 4, 4)), ((4, 6, 4, 4), (4, 6, 4, 4)),
                                           ((2, 4, 1, 1), (2, 4, 3, 5))])
def test_gc_backward_matches_autograd(gc_pair, yshape, sshape):
    ours, _ = gc_pair
    ours.train()
    refd = R.GaussianConditional(None).double().train()
    torch.manual_seed(7)
    C = sshape[1]
    scales = torch.exp(torch.empty(*sshape).uniform_(np.log(0.05), np.log(30)))
    y, noise = torch.randn(*yshape) * 3, torch.rand(*yshape) - 0.5
    g_lik, g_out, w = torch.randn(*sshape), torch.randn(*yshape), torch.randn(C)
    yd, sd = y.to(DEV).requires_grad_(True), scales.to(DEV).requires_grad_(True)
    yh, yl = ours(yd, sd, noise=noise.to(DEV))
    ((yl * g_lik.to(DEV)).sum() + (yh * g_out.to(DEV)).sum() + (ours.last_log_likelihood_sums * w.to(DEV)).sum()).backward()
    y64, s64 = y.double().requires_grad_(True), scales.double().requires_grad_(True)
    yh_r, yl_r = refd(y64, s64, noise=noise.double())
    ln_r = torch.log(yl_r).transpose(0, 1).reshape(C, -1).sum(1)
    ((yl_r * g_lik.double()).sum() + (yh_r * g_out.double()).sum() + (ln_r * w.double()).sum()).backward()
    assert torch.allclose(yd.grad.cpu().double(), y64.grad, rtol=2e-3, atol=2e-5)
    assert torch.allclose(sd.grad.cpu().double(), s64.grad, rtol=2e-3, atol=2e-5)


def test_lnsum_op_forward_backward():
    torch.manual_seed(8)
    lik = (torch.rand(5, 7, 3, 3) * 0.9 + 1e-6)
    ld = lik.to(DEV).requires_grad_(True)
    s = mm.ops.channel_log_likelihood_sums(ld)
    want = torch.log(lik.double()).transpose(0, 1).reshape(7, -1).sum(1)
    assert torch.allclose(s.cpu().double(), want, rtol=1e-5)
    w = torch.randn(7)
    (s * w.to(DEV)).sum().backward()
    assert torch.allclose(ld.grad.cpu(), (w.view(1, 7, 1, 1) / lik), rtol=1e-5)


# ------------------------------------------------------------------------------------------------ a8
GDN_SHAPES = [(2, 10, 6, 5), (1, 1, 16, 16), (2, 3, 32, 32), (3, 16, 8, 8), (2, 50, 16, 16), (1, 100, 8, 16),
              (2, 33, 4, 4), (4, 300, 1, 1), (1, 128, 12, 12), (2, 17, 7, 9), (2, 3, 5, 7), (3, 4, 8, 8), (2, 2, 64, 64),
              (5, 1, 3, 3)]


def _pair_gdn(C, inverse, seed=0, precision="fp32"):
    g = torch.Generator().manual_seed(seed)
    ref = R.GDN(C, inverse=inverse)
    with torch.no_grad():
        ref.gamma.add_(torch.rand(C, C, generator=g) * 0.05)
        ref.beta.add_(torch.rand(C, generator=g) * 0.5)
    ours = mm.GDN(C, inverse=inverse, precision=precision)
    ours.load_state_dict(ref.state_dict())
    return ours.to(DEV), ref


@pytest.mark.parametrize("shape", GDN_SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_forward_fp32(shape, inverse):
    """fp32 SIMT contraction: tolerance = fp32 summation-order noise over C terms (rtol 2e-5)."""
    torch.manual_seed(9)
    ours, ref = _pair_gdn(shape[1], inverse)
    x = torch.randn(*shape)
    y = ours(x.to(DEV))
    assert torch.allclose(y.cpu(), ref(x), rtol=2e-5, atol=1e-6)


TC_SHAPES = [(2, 50, 64, 64), (1, 100, 64, 64), (3, 16, 48, 40), (4, 128, 32, 32), (2, 33, 50, 50), (1, 64, 64, 65),
             (1, 200, 64, 64), (2, 24, 37, 59), (1, 160, 64, 64)]


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("precision,rtol", [("tf32", 1e-3), ("3xtf32", 2e-5)])
def test_gdn_forward_tensor_core(shape, inverse, precision, rtol):
    """tcgen05 contraction.  Stated tolerances: single-pass TF32 rounds x^2 and gamma to 10-bit mantissas
    (2^-11 = 4.9e-4 relative per term, halved by the square root) -> rtol 1e-3 on y; the three-pass hi/lo split
    is held to the same 2e-5 as the fp32 SIMT kernel."""
    torch.manual_seed(15)
    ours, ref = _pair_gdn(shape[1], inverse, precision=precision)
    x = torch.randn(*shape)
    before = mm.launch_count()
    y = ours(x.to(DEV))
    torch.cuda.synchronize()
    wide = precision == "tf32" and shape[1] > 128  # 129 .. 256 channels: pack gamma + the streamed-operand kernel
    assert mm.launch_count() == before + (2 if wide else 1), "re-parametrisation is fused: one launch per GDN forward"
    if precision == "3xtf32" and shape[1] > 160:
        rtol = 2e-5  # too many TMEM columns for the split: dispatches to the fp32 SIMT kernel, same tolerance
    want = ref(x)
    assert torch.allclose(y.detach().cpu(), want, rtol=rtol, atol=rtol * 0.1), (y.detach().cpu() - want).abs().max()
    y32 = mm.ops.gdn(x.to(DEV), ours.beta_reparam(ours.beta), ours.gamma_reparam(ours.gamma), inverse, "fp32")
    assert torch.allclose(y, y32, rtol=rtol, atol=rtol * 0.1)


def test_gdn_golden():
    fx = load_golden()
    for tag, inv in (("gdn", False), ("igdn", True)):
        ours = mm.GDN(10, inverse=inv, precision="fp32")
        with torch.no_grad():
            ours.beta.copy_(torch.from_numpy(fx[f"{tag}_beta"]))
            ours.gamma.copy_(torch.from_numpy(fx[f"{tag}_gamma"]))
        y = ours.to(DEV)(torch.from_numpy(fx[f"{tag}_x"]).to(DEV))
        assert np.allclose(y.detach().cpu().numpy(), fx[f"{tag}_y"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 10, 6, 5), (2, 3, 16, 16), (2, 50, 8, 8), (3, 33, 4, 4), (2, 128, 4, 4),
                                   (5, 20, 1, 1), (1, 70, 9, 9), (2, 3, 5, 7), (3, 4, 8, 8), (2, 1, 64, 64), (4, 2, 33, 1)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_fp32(shape, inverse):
    torch.manual_seed(10)
    C = shape[1]
    ours, ref = _pair_gdn(C, inverse)
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    xd = x.to(DEV).requires_grad_(True)
    (ours(xd) * g.to(DEV)).sum().backward()
    x64 = x.double().requires_grad_(True)
    (refd(x64) * g.double()).sum().backward()
    assert torch.allclose(xd.grad.cpu().double(), x64.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(ours.beta.grad.cpu().double(), refd.beta.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(ours.gamma.grad.cpu().double(), refd.gamma.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("shape", [(2, 50, 64, 64), (1, 100, 64, 64), (3, 16, 48, 40), (4, 128, 32, 32), (2, 33, 50, 50),
                                   (1, 64, 64, 65), (2, 24, 37, 59), (2, 120, 48, 48)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_tensor_core(shape, inverse):
    """Fused tcgen05 backward (three single-pass TF32 contractions).  Stated tolerance: TF32 rounds x^2, u and gamma
    to 10-bit mantissas -> 2e-3 relative (of the largest entry) on dx; d gamma / d beta are sums over all pixels
    of tf32-rounded products, held to 2e-3 of their largest entry."""
    torch.manual_seed(16)
    C = shape[1]
    ours, ref = _pair_gdn(C, inverse, precision="tf32")
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    xd = x.to(DEV).requires_grad_(True)
    before = mm.launch_count()
    (ours(xd) * g.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    if C <= 112:  # larger C: gamma and gamma^T no longer fit shared memory next to the pixel-major operands -> SIMT
        assert mm.launch_count() - before == 1 + 2, "fwd: 1 fused kernel; bwd: fused kernel + partial reduce"
    x64 = x.double().requires_grad_(True)
    (refd(x64) * g.double()).sum().backward()

    def close(a, b, tol):
        return ((a.cpu().double() - b).abs().max() / b.abs().max()).item() <= tol

    assert close(xd.grad, x64.grad, 2e-3), ((xd.grad.cpu().double() - x64.grad).abs().max(), x64.grad.abs().max())
    assert close(ours.beta.grad, refd.beta.grad, 2e-3)
    assert close(ours.gamma.grad, refd.gamma.grad, 2e-3)




# (KP8 -> groups, stages) instances of gdn_tc_fwd2.cu: C=16/20 -> 4 groups, 33/40 -> 3, 50/64 -> 2, 100/104 -> 1 x 3 stages,
# 128 -> 1 x 2 stages; C % 8 == 0 takes beta from shared memory, otherwise through the constant MMA column
@pytest.mark.parametrize("shape", [(10, 50, 64, 64), (5, 100, 128, 64), (10, 64, 64, 64), (12, 20, 64, 64),
                                   (10, 40, 32, 128), (10, 33, 64, 64), (10, 104, 64, 64), (10, 128, 64, 64),
                                   (20, 16, 16, 128), (300, 16, 8, 16), (1, 50, 128, 128), (3, 112, 16, 8),
                                   (2, 77, 16, 16)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_forward_tensor_core_tma(shape, inverse):
    """TMA-in / TMA-out tcgen05 forward (gdn_tc_fwd2.cu) against the oracle; single-pass TF32 tolerance 1e-3."""
    torch.manual_seed(19)
    ours, ref = _pair_gdn(shape[1], inverse, precision="tf32")
    x = torch.randn(*shape)
    xd = x.to(DEV)
    before = mm.launch_count()
    y = ours(xd)
    torch.cuda.synchronize()
    assert mm.launch_count() == before + 1
    want = ref(x)
    assert torch.allclose(y.detach().cpu(), want, rtol=1e-3, atol=1e-4), (y.detach().cpu() - want).abs().max()
    assert torch.equal(ours(xd[1:3].contiguous()), y[1:3]), "batch independence"
    assert torch.equal(xd.cpu(), x), "the input must not be modified (y is written in place in shared memory only)"


def _gdn_bwd_variant(x, g, precision="tf32"):
    B, C = x.shape[:2]
    HW = x.numel() // (B * C)
    return int(mm._lib.lib().mmnc_gdn_backward_variant(x.data_ptr(), g.data_ptr(), B, C, HW,
                                                        mm.ops.GDN_PRECISION[precision]))


# (P, groups, stages) instances of gdn_tc_bwd2.cu: C=20 -> (32,2,3), 40 -> (48,2,3), 50 / 63 -> (64,2,3), 64 -> (80,2,2),
# 90 -> (96,1,1), 100 / 110 -> (112,1,1); 320-640 tiles on 148 CTAs: groups with 1, 2 and 0 tiles all occur
@pytest.mark.parametrize("shape", [(10, 50, 64, 64), (5, 100, 128, 64), (10, 64, 64, 64), (12, 20, 64, 64),
                                   (10, 40, 32, 128), (10, 90, 64, 64), (10, 63, 64, 64), (10, 110, 64, 64),
                                   (20, 16, 16, 128), (300, 16, 8, 16), (1, 50, 128, 128), (3, 111, 16, 8),
                                   (2, 77, 16, 16)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_tensor_core_pipelined(shape, inverse):
    """TMA-fed, software-pipelined tcgen05 backward (gdn_tc_bwd2.cu) against the float64 oracle; same stated
    tolerance as the first-generation fused kernel (2e-3 of the largest entry)."""
    torch.manual_seed(17)
    C = shape[1]
    ours, ref = _pair_gdn(C, inverse, precision="tf32")
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    xd, gd = x.to(DEV).requires_grad_(True), g.to(DEV)
    assert _gdn_bwd_variant(xd, gd) == 3, "shape should take the pipelined kernel"
    before = mm.launch_count()
    ours(xd).backward(gd)
    torch.cuda.synchronize()
    assert mm.launch_count() - before == 1 + 2
    x64 = x.double().requires_grad_(True)
    refd(x64).backward(g.double())

    def close(a, b, tol):
        return ((a.cpu().double() - b).abs().max() / b.abs().max()).item() <= tol

    assert close(xd.grad, x64.grad, 2e-3), ((xd.grad.cpu().double() - x64.grad).abs().max(), x64.grad.abs().max())
    assert close(ours.beta.grad, refd.beta.grad, 2e-3)
    assert close(ours.gamma.grad, refd.gamma.grad, 2e-3)


@pytest.mark.parametrize("shape", [(6, 128, 64, 64), (3, 128, 128, 128), (10, 112, 32, 32), (5, 120, 64, 32),
                                   (4, 127, 32, 64), (300, 128, 8, 16), (1, 128, 16, 8)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_tensor_core_streamed_gamma(shape, inverse):
    """112 <= C <= 128 (BASELINE configs C3 / C4: GDN(128)): the TMA-fed fused backward with the gamma operand streamed
    through one shared-memory buffer (variant 4), against the float64 oracle at the TF32 bar (2e-3 of the largest
    entry), deterministic run to run; 96-600 tiles on 148 CTAs, so CTAs with 0, 1 and several tiles all occur."""
    torch.manual_seed(19)
    C = shape[1]
    ours, ref = _pair_gdn(C, inverse, precision="tf32")
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    xd, gd = x.to(DEV).requires_grad_(True), g.to(DEV)
    assert _gdn_bwd_variant(xd, gd) == 4, "shape should take the streamed-gamma kernel"
    before = mm.launch_count()
    ours(xd).backward(gd)
    torch.cuda.synchronize()
    assert mm.launch_count() - before == 1 + 3  # forward; pack gamma tiles, fused backward, partial reduce
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
    assert all(torch.equal(a, b) for a, b in zip(first, (xd.grad, ours.beta.grad, ours.gamma.grad)))


@pytest.mark.parametrize("shape", [(32, 50, 128, 128), (16, 100, 128, 128), (24, 64, 128, 128)])
def test_gdn_backward_pipelined_long_sequences(shape):
    """Dozens of tiles per CTA (every stage and mbarrier phase wraps many times): the pipelined kernel against the
    fp32 SIMT kernel on the same device, and bit-identical results run to run (fixed-order partial reduction)."""
    torch.manual_seed(18)
    C = shape[1]
    ours, _ = _pair_gdn(C, False, precision="tf32")
    x = torch.randn(*shape, device=DEV)
    g = torch.randn(*shape, device=DEV)
    # C = 100 with >= 3 tiles per CTA takes the x-prefetch variant (5): second x buffer + streamed gamma
    assert _gdn_bwd_variant(x, g) == (5 if C == 100 else 3)
    beta, gamma = ours.beta_reparam(ours.beta).detach(), ours.gamma_reparam(ours.gamma).detach()
    outs = []
    for prec in ("tf32", "tf32", "fp32"):
        xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
        outs.append(torch.autograd.grad(mm.ops.gdn(xr, br, gr, False, prec), [xr, br, gr], g))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b), "pipelined backward is not deterministic"
    with torch.no_grad():  # the forward of the same layers (dozens of tiles per compute group) against fp32
        y_tc, y_32 = mm.ops.gdn(x, beta, gamma, False, "tf32"), mm.ops.gdn(x, beta, gamma, False, "fp32")
    assert torch.allclose(y_tc, y_32, rtol=1e-3, atol=1e-4)
    for a, b in zip(outs[0], outs[2]):
        assert ((a - b).abs().max() / b.abs().max()).item() <= 2e-3



def test_gdn_tma_kernels_fall_back_on_unaligned_or_odd_shapes():
    """The TMA-fed kernels need 16-byte aligned tensors and H*W % 128 == 0; anything else must silently take the
    first-generation kernels and give the same numbers."""
    torch.manual_seed(23)
    C, B, H, W = 50, 10, 64, 64
    ours, _ = _pair_gdn(C, False, precision="tf32")
    x_al = torch.randn(B, C, H, W, device=DEV)
    x_un = torch.empty(B * C * H * W + 1, device=DEV)[1:].view(B, C, H, W)  # data pointer off by 4 bytes
    x_un.copy_(x_al)
    g = torch.randn(B, C, H, W, device=DEV)
    assert x_un.data_ptr() % 16 != 0 and x_un.is_contiguous()
    assert _gdn_bwd_variant(x_al, g) == 3 and _gdn_bwd_variant(x_un, g) == 2
    outs = []
    for x in (x_al, x_un):
        xr = x.detach().requires_grad_(True)
        y = ours(xr)
        (gx,) = torch.autograd.grad(y, [xr], g)
        outs.append((y.detach(), gx))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-6), "forward generations disagree"
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-5), "backward generations disagree"
    odd = torch.randn(3, C, 30, 31, device=DEV)  # H*W = 930: no 128-pixel TMA tiles
    assert _gdn_bwd_variant(odd, odd) == 2



@pytest.mark.parametrize("shape", [(16, 50, 256, 256), (16, 100, 128, 128)])
def test_gdn_tensor_core_full_size_properties(shape):
    """The two largest layer shapes of BASELINE config C2 (batch 16 of its 64) through the TMA-fed tensor-core kernels:
    size-independent properties, and torch's own fp32 ops on the same device as the reference."""
    torch.manual_seed(29)
    B, C, H, W = shape
    ours, _ = _pair_gdn(C, False, precision="tf32")
    beta, gamma = ours.beta_reparam(ours.beta).detach(), ours.gamma_reparam(ours.gamma).detach()
    x = torch.randn(*shape, device=DEV)
    g = torch.randn(*shape, device=DEV)
    assert _gdn_bwd_variant(x, g) == (5 if C == 100 else 3)

    def run(xi, gi):
        xr, br, gr = xi.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
        y = mm.ops.gdn(xr, br, gr, False, "tf32")
        return (y.detach(),) + torch.autograd.grad(y, [xr, br, gr], gi)

    y, dx, db, dg = run(x, g)
    # (1) odd symmetry in x and in g, batch independence: exact (the same roundings happen in the same order)
    y2, dx2, db2, dg2 = run(-x, g)
    assert torch.equal(y2, -y) and torch.equal(dx2, dx) and torch.equal(db2, -db) and torch.equal(dg2, -dg)
    y3, dx3, _, _ = run(x[3:5].contiguous(), g[3:5].contiguous())
    assert torch.equal(y3, y[3:5]) and torch.equal(dx3, dx[3:5])
    # (2) d gamma / d beta are sums over images: two halves of the batch add up to the whole
    _, _, dba, dga = run(x[: B // 2].contiguous(), g[: B // 2].contiguous())
    _, _, dbb, dgb = run(x[B // 2:].contiguous(), g[B // 2:].contiguous())
    assert ((dga + dgb - dg).abs().max() / dg.abs().max()).item() < 1e-5
    assert ((dba + dbb - db).abs().max() / db.abs().max()).item() < 1e-5
    # (3) torch's fp32 ops (conv2d with TF32 off) on the same device
    xr, br, gr = x.clone().requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        want = xr * torch.rsqrt(torch.nn.functional.conv2d(xr * xr, gr.reshape(C, C, 1, 1), br))
        wdx, wdb, wdg = torch.autograd.grad(want, [xr, br, gr], g)
    assert torch.allclose(y, want.detach(), rtol=1e-3, atol=1e-4)
    for got, ref in ((dx, wdx), (db, wdb), (dg, wdg)):
        assert ((got - ref).abs().max() / ref.abs().max()).item() <= 2e-3


def test_gdn_reparam_lower_bound_gradient():
    """A.2 / A.5: below the bound the gradient passes only if it is negative."""
    C = 4
    ours, ref = _pair_gdn(C, False)
    with torch.no_grad():
        for m in (ours, ref):
            m.beta[:2] = 1e-4  # below bound sqrt(1e-6 + 2^-36)
            m.gamma[0, :] = -1.0
    x, g = torch.randn(2, C, 3, 3), torch.randn(2, C, 3, 3)
    (ours(x.to(DEV)) * g.to(DEV)).sum().backward()
    (ref(x) * g).sum().backward()
    assert torch.allclose(ours.beta.grad.cpu(), ref.beta.grad, rtol=1e-3, atol=1e-5)
    assert torch.allclose(ours.gamma.grad.cpu(), ref.gamma.grad, rtol=1e-3, atol=1e-5)


def test_gdn_full_size_properties():
    """BASELINE config C2's largest layer (GDN(50) on 256^2) at batch 8: size-independent properties."""
    torch.manual_seed(11)
    C, B, H = 50, 8, 256
    ours, _ = _pair_gdn(C, False)
    x = torch.randn(B, C, H, H, device=DEV)
    y = ours(x)
    # (1) sample check against torch ops on the same device
    beta, gamma = ours.beta_reparam(ours.beta), ours.gamma_reparam(ours.gamma)
    sl = (slice(0, 2), slice(None), slice(100, 110), slice(None))
    with torch.backends.cudnn.flags(allow_tf32=False):  # torch's own conv would otherwise run in TF32
        want = x[sl] * torch.rsqrt(torch.nn.functional.conv2d(x[sl] ** 2, gamma.reshape(C, C, 1, 1), beta))
    assert torch.allclose(y[sl], want, rtol=2e-5, atol=1e-6)
    # (2) odd symmetry and (3) batch independence
    assert torch.equal(ours(-x), -y)
    assert torch.equal(ours(x[3:5]), y[3:5])
    # (4) scaling law: y(s x; s^2 beta, gamma) = y(x; beta, gamma)
    y2 = mm.ops.gdn(3.0 * x, 9.0 * beta, gamma, False, "fp32")
    assert torch.allclose(y2, y, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ a10-a12
def test_build_indexes_bit_exact(gc_pair):
    ours, ref = gc_pair
    table = R.get_scale_table()
    scales = torch.cat([torch.exp(torch.empty(20000).uniform_(np.log(0.01), np.log(400))), table,
                        table * (1 + 1e-7), table * (1 - 1e-7),
                        torch.tensor([0.0, -1.0, float("inf"), float("nan"), 0.11, 256.0])]).reshape(1, 1, -1)
    assert torch.equal(ours.build_indexes(scales.to(DEV)).cpu(), ref.build_indexes(scales))
    fx = load_golden()
    assert np.array_equal(ours.build_indexes(torch.from_numpy(fx["gc_scales"]).to(DEV)).cpu().numpy(), fx["gc_idx"])


def test_rans_golden_streams(gc_pair):
    ours, _ = gc_pair
    fx = load_golden()
    for sym, idx, want in zip(fx["rans_sym"], fx["rans_idx"], fx["rans_bytes"]):
        s = mm.ops.rans_encode(torch.from_numpy(sym).to(DEV).reshape(1, -1), torch.from_numpy(idx).to(DEV).reshape(1, -1),
                               0, ours._rans_tables())
        assert s == [want.tobytes()]
        back = mm.ops.rans_decode(s, torch.from_numpy(idx).to(DEV).reshape(1, -1), 0, sym.size, ours._rans_tables())
        assert np.array_equal(back.cpu().numpy().reshape(-1), sym)


@pytest.mark.parametrize("B,shape", [(1, (8, 1, 1)), (7, (128, 1, 1)), (64, (300, 1, 1)), (5, (12, 4, 4)),
                                     (33, (3, 5, 7))])
def test_gc_compress_bit_exact(gc_pair, B, shape):
    ours, ref = gc_pair
    torch.manual_seed(12)
    scales = torch.exp(torch.empty(B, *shape).uniform_(np.log(0.05), np.log(64)))
    y = torch.randn(B, *shape) * scales
    esc = torch.rand(B, *shape) < 0.01  # 1 % outliers -> bypass coding
    y[esc] = y[esc] * 40 + 300
    idx_r = ref.build_indexes(scales)
    idx = ours.build_indexes(scales.to(DEV))
    assert torch.equal(idx.cpu(), idx_r)
    ref.marshalling = "lean"
    s_r = ref.compress(y, idx_r)
    s = ours.compress(y.to(DEV), idx)
    assert s == s_r
    assert torch.equal(ours.decompress(s, idx).cpu(), ref.decompress(s_r, idx_r))
    assert torch.equal(ours.decompress(s, idx).cpu(), torch.round(y))


@pytest.mark.parametrize("B,C,hw", [(1, 4, (1, 1)), (64, 300, (1, 1)), (9, 32, (2, 3)), (256, 512, (1, 1))])
def test_eb_compress_bit_exact(B, C, hw):
    torch.manual_seed(13)
    ours, ref = _pair_eb(C, True)
    ours.update(), ref.update()
    assert torch.equal(ours._quantized_cdf, ref._quantized_cdf)
    ours.to(DEV)
    z = torch.randn(B, C, *hw) * 6
    z[0, 0] = 90.0
    ref.marshalling = "lean"
    s_r = ref.compress(z)
    s = ours.compress(z.to(DEV))
    assert s == s_r
    zh = ours.decompress(s, hw)
    assert torch.equal(zh.cpu(), ref.decompress(s_r, hw))
    ours.eval()
    assert torch.equal(zh, ours(z.to(DEV))[0]), "decompress(compress(z)) == eval-mode quantisation"


def test_rans_error_paths(gc_pair):
    ours, _ = gc_pair
    with pytest.raises(ValueError, match="same size"):
        ours.compress(torch.zeros(2, 4, 1, 1, device=DEV), torch.zeros(2, 4, 4, 4, dtype=torch.int32, device=DEV))
    bad_idx = torch.full((1, 4, 1, 1), 99, dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError, match="malformed"):
        ours.compress(torch.zeros(1, 4, 1, 1, device=DEV), bad_idx)
    idx = torch.zeros(1, 64, 1, 1, dtype=torch.int32, device=DEV)
    s = ours.compress(torch.zeros(1, 64, 1, 1, device=DEV), idx)
    with pytest.raises(ValueError, match="corrupt|truncated"):
        ours.decompress([s[0][:4]], idx)


def test_rans_full_size_roundtrip(gc_pair):
    """BASELINE config 5 upper end: 1024 images, and a shape-consistent 192 x 16 x 16 latent (49 152 symbols)."""
    ours, _ = gc_pair
    torch.manual_seed(14)
    for B, shape in ((1024, (128, 1, 1)), (16, (192, 16, 16))):
        scales = torch.exp(torch.empty(B, *shape, device=DEV).uniform_(np.log(0.05), np.log(64)))
        y = torch.randn(B, *shape, device=DEV) * scales
        idx = ours.build_indexes(scales)
        s = ours.compress(y, idx)
        assert torch.equal(ours.decompress(s, idx), torch.round(y))
        ours.eval()
        _, lik = ours(y, scales)
        est_bits = float(-torch.log2(lik).sum())
        actual_bits = 8 * sum(len(b) for b in s)
        assert actual_bits >= est_bits * 0.95 and actual_bits <= est_bits * 1.10 + 64 * B, (actual_bits, est_bits)
        # spot-check a few streams against the CPU oracle
        cdf, ln, off = (t.cpu().numpy() for t in (ours._quantized_cdf, ours._cdf_length, ours._offset))
        for i in (0, B // 2, B - 1):
            want = native.encode_with_indexes_np(torch.round(y[i]).int().cpu().numpy(), idx[i].cpu().numpy(), cdf, ln, off)
            assert s[i] == want


# ------------------------------------------------------------------------------------------------ a1, a6, a7: wrappers
CONFIGS = [(1, ("mono",), 8, 8), (2, ("rgb", "depth_euclidean", "normal", "semantic"), 12, 8),
           (3, ("rgb", "depth_euclidean", "normal"), 14, 12), (4, ("rgb", "depth_euclidean", "normal", "semantic"), 13, 8)]


@pytest.mark.parametrize("kind,tasks,l,c", CONFIGS)
def test_wrapper_loss_parity_from_same_likelihoods(kind, tasks, l, c):
    """RD-loss formulas in isolation: feed both implementations the SAME likelihoods and reconstructions."""
    torch.manual_seed(20 + kind)
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2)
    ref = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2)
    ref.load_state_dict(ours.state_dict())
    if kind != 1:
        with torch.no_grad():
            lv = torch.randn(len(tasks)) * 0.3
            ours.loss_balancer.log_vars.copy_(lv), ref.loss_balancer.log_vars.copy_(lv)
    ours.to(DEV)
    B, M, N = 2, ours.model["compressor"].M, c * len(tasks)
    lik = {"y": torch.rand(B, M, 4, 4) * 0.9 + 1e-7, "z": torch.rand(B, N, 1, 1) * 0.9 + 1e-7}
    batch = mm.synthetic_batch(tasks, B, size=32, seed=3)
    x_hats = {t: torch.randn(B, mm.task_parameters[t]["out_channels"], 32, 32) for t in tasks}
    lik_d = {k: v.to(DEV).requires_grad_(True) for k, v in lik.items()}
    xh_d = {k: v.to(DEV).requires_grad_(True) for k, v in x_hats.items()}
    loss, logs = ours.rate_distortion_loss({k: v.to(DEV) for k, v in batch.items()}, xh_d, lik_d, "val")
    lik_r = {k: v.clone().requires_grad_(True) for k, v in lik.items()}
    xh_r = {k: v.clone().requires_grad_(True) for k, v in x_hats.items()}
    rec, l1 = ref.multitask_reconstruction_loss(batch, xh_r, "val")
    comp, l2 = ref.multitask_compression_loss(lik_r, xh_r, "val")
    want = ref.lmbda * rec + comp
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    want_logs = {"val/rec_loss": rec, "val/compression_loss": comp, "val/loss": want, **l1, **l2}
    assert set(logs) == set(want_logs)
    for k, v in want_logs.items():
        assert abs(float(logs[k]) - float(v)) <= 1e-5 * abs(float(v)) + 1e-7, k
    loss.backward(), want.backward()
    for k in lik:
        assert torch.allclose(lik_d[k].grad.cpu(), lik_r[k].grad, rtol=1e-4, atol=1e-9), k
    for k in x_hats:
        assert torch.allclose(xh_d[k].grad.cpu(), xh_r[k].grad, rtol=1e-4, atol=1e-8), k
    if kind != 1:
        assert torch.allclose(ours.loss_balancer.log_vars.grad.cpu(), ref.loss_balancer.log_vars.grad, rtol=1e-4)
    # reference-shaped entry points agree with the fused one
    c2, _ = ours.multitask_compression_loss(lik_d, xh_d, "val")
    r2, _ = ours.multitask_reconstruction_loss({k: v.to(DEV) for k, v in batch.items()}, xh_d, "val")
    assert abs(c2.item() - comp.item()) <= 1e-5 * abs(comp.item()) and abs(r2.item() - rec.item()) <= 1e-5 * abs(rec.item())


@pytest.mark.parametrize("kind,tasks,l,c", CONFIGS)
def test_wrapper_end_to_end_vs_oracle(kind, tasks, l, c):
    """Whole model at 256^2 (the reference's y (B,M,1,1) / scales (B,M,4,4) broadcast included).  Convolutions run
    on different devices (cuDNN vs CPU) in the two arms, hence the looser tolerance on the scalar loss."""
    torch.manual_seed(30 + kind)
    ours = mm.build_compressor(kind, tasks, l, c, lmbda=1e-2)
    ref = orm.ReferenceCompressor(kind, tasks, l, c, lmbda=1e-2).eval()
    ref.load_state_dict(ours.state_dict())
    ours.to(DEV).eval()
    batch = mm.synthetic_batch(tasks, 2, size=256, seed=21)
    bd = {k: v.to(DEV) for k, v in batch.items()}
    with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        x_hats, lik = ours(bd)
        loss, logs = ours.rate_distortion_loss(bd, x_hats, lik, "val")
        want, want_logs = ref.rd_loss(batch, "val")
    assert lik["y"].shape == (2, ours.model["compressor"].M, 4, 4) and lik["z"].shape[2:] == (1, 1)
    assert all(x_hats[t].shape[-2:] == (256, 256) for t in tasks)
    assert abs(loss.item() - want.item()) <= 5e-4 * abs(want.item())
    for k in want_logs:
        assert abs(float(logs[k]) - float(want_logs[k])) <= 2e-3 * abs(float(want_logs[k])) + 1e-6, k


def test_train_step_gradients_vs_oracle():
    """One -m 3 train step: gradients of a sample of parameters against the oracle's autograd (CPU)."""
    torch.manual_seed(41)
    tasks = ("rgb", "depth_euclidean", "normal")
    ours = mm.build_compressor(3, tasks, 14, 12, lmbda=1e-2)
    ref = orm.ReferenceCompressor(3, tasks, 14, 12, lmbda=1e-2).train()
    ref.load_state_dict(ours.state_dict())
    ours.to(DEV).train()
    batch = mm.synthetic_batch(tasks, 2, size=256, seed=21)
    bd = {k: v.to(DEV) for k, v in batch.items()}
    nz = torch.rand(2, 36, 1, 1) - 0.5
    ny = torch.rand(2, 14, 1, 1) - 0.5
    # inject the same noise into both arms through the modules' test hook
    co, cr = ours.model["compressor"], ref.model["compressor"]
    eb_f, gc_f, eb_rf, gc_rf = (co.entropy_bottleneck.forward, co.gaussian_conditional.forward,
                                cr.entropy_bottleneck.forward, cr.gaussian_conditional.forward)
    co.entropy_bottleneck.forward = lambda x, training=None: eb_f(x, training, noise=nz.to(DEV))
    co.gaussian_conditional.forward = lambda y, s, means=None, training=None: gc_f(y, s, means, training, noise=ny.to(DEV))
    cr.entropy_bottleneck.forward = lambda x, training=None: eb_rf(x, training, noise=nz)
    cr.gaussian_conditional.forward = lambda y, s, means=None, training=None: gc_rf(y, s, means, training, noise=ny)
    with torch.backends.cudnn.flags(allow_tf32=False):
        x_hats, lik = ours(bd)
        loss, _ = ours.rate_distortion_loss(bd, x_hats, lik, "train")
        loss.backward()
    want, _ = ref.rd_loss(batch, "train")
    want.backward()
    assert abs(loss.item() - want.item()) <= 5e-4 * abs(want.item())
    ref_params = dict(ref.named_parameters())
    checked = 0
    for name, p in ours.named_parameters():
        if p.grad is None or name.endswith("quantiles"):
            continue
        g, w = p.grad.cpu(), ref_params[name].grad
        denom = w.abs().max().item()
        if denom < 1e-12:
            continue
        err = (g - w).abs().max().item() / denom
        assert err < 2e-2, (name, err)
        checked += 1
    assert checked > 100


def test_model_compress_decompress_roundtrip_and_module_parity():
    """-m 2 at reduced resolution where y and scales have the same shape, so compress() is well defined
    (at 256^2 the reference raises, Appendix B1 — checked below)."""
    torch.manual_seed(42)
    tasks = ("rgb", "depth_euclidean")
    ours = mm.build_compressor(2, tasks, 12, 8, lmbda=1e-2)
    ours.update_bottleneck_values()
    ours.to(DEV).eval()
    c = ours.model["compressor"]
    x = torch.randn(3, 16, 64, 64, device=DEV)  # backbone input: y is 4x4, z is 1x1, scales 4x4
    with torch.no_grad():
        out = c.compress(x)
        y = c.g_a(x)
        z = c.h_a(torch.abs(y))
    # module-level parity with the oracle on the SAME y, z
    ref_eb, ref_gc = R.EntropyBottleneck(16), R.GaussianConditional(None)
    ref_eb.load_state_dict({k: v.cpu() for k, v in c.entropy_bottleneck.state_dict().items()})
    ref_gc.update_scale_table(R.get_scale_table())
    ref_eb.marshalling = ref_gc.marshalling = "lean"
    assert out["strings"][1] == ref_eb.compress(z.cpu())
    with torch.no_grad():
        z_hat = c.entropy_bottleneck.decompress(out["strings"][1], z.shape[-2:])
        idx = c.gaussian_conditional.build_indexes(c.h_s(z_hat))
    assert out["strings"][0] == ref_gc.compress(y.cpu(), idx.cpu())
    with torch.no_grad():
        rec = c.decompress(out["strings"], out["shape"])["x_hat"]
        y_hat = c.gaussian_conditional.decompress(out["strings"][0], idx, z_hat.dtype)
    assert torch.equal(y_hat, torch.round(y)) and rec.shape == x.shape
    batch = mm.synthetic_batch(tasks, 1, size=256, seed=21, device=DEV)
    with pytest.raises(ValueError, match="same size"):
        ours.compress(batch)  # Appendix B1: y (B,M,1,1) vs indexes (B,M,4,4)


def test_decompress_wrapper_matches_forward():
    """MultiTaskCompressor.decompress (mtc.py:536-549) reproduces the eval forward's reconstruction."""
    torch.manual_seed(43)
    tasks = ("rgb", "depth_euclidean", "normal")
    ours = mm.build_compressor(3, tasks, 15, 12, lmbda=1e-2)
    ours.update_bottleneck_values()
    ours.to(DEV).eval()
    c = ours.model["compressor"]
    with torch.no_grad():
        z = torch.randn(2, 36, 1, 1, device=DEV) * 3
        z_strings = c.entropy_bottleneck.compress(z)
        z_hat = c.entropy_bottleneck.decompress(z_strings, (1, 1))
        scales = c.h_s(z_hat)
        idx = c.gaussian_conditional.build_indexes(scales)
        y = torch.randn(2, 15, 4, 4, device=DEV) * scales
        y_strings = c.gaussian_conditional.compress(y, idx)
        x_hats = ours.decompress([y_strings, z_strings], (1, 1))
        want = ours.forward_output_heads(torch.round(y))
    # cuDNN's transposed convolutions are not run-to-run bit-reproducible: compare numerically
    assert all(torch.allclose(x_hats[t], want[t], rtol=1e-4, atol=1e-5) for t in tasks)


def test_launch_counter_moves():
    before = mm.launch_count()
    mm.GDN(4).to(DEV)(torch.randn(1, 4, 8, 8, device=DEV))
    assert mm.launch_count() == before + 1  # re-parametrisation fused into the contraction


# ------------------------------------------------------------------------------------------------ (f1) channels-last
@pytest.mark.parametrize("shape", [(4, 50, 64, 64), (2, 100, 32, 32), (3, 64, 16, 16), (2, 3, 64, 64), (5, 33, 8, 8),
                                   (2, 111, 16, 16), (2, 128, 32, 32), (1, 17, 64, 64), (2, 20, 30, 31), (3, 3, 5, 7),
                                   (2, 300, 4, 4), (3, 1, 16, 16), (2, 256, 64, 32), (1, 192, 64, 64)])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_channels_last_matches_oracle(shape, inverse):
    """NHWC tensors through GDN (native kernels where they exist - C <= 4 streaming, 16 <= C <= 128 tensor cores with
    row access, vector widths 4 / 2 / 1 - and converted otherwise): output and input gradient keep the channels-last
    format and agree with the float64 oracle at the TF32 bars (1e-3 forward, 2e-3 of the largest entry backward)."""
    torch.manual_seed(31)
    C = shape[1]
    ours, ref = _pair_gdn(C, inverse, precision="auto")
    refd = R.GDN(C, inverse=inverse).double()
    refd.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x, g = torch.randn(*shape), torch.randn(*shape)
    cl = torch.channels_last
    xd = x.to(DEV).contiguous(memory_format=cl).requires_grad_(True)
    y = ours(xd)
    y.backward(g.to(DEV).contiguous(memory_format=cl))
    x64 = x.double().requires_grad_(True)
    y64 = refd(x64)
    y64.backward(g.double())
    assert y.shape == x.shape and y.is_contiguous(memory_format=cl) and xd.grad.is_contiguous(memory_format=cl)

    def close(a, b, tol):
        return ((a.detach().cpu().double() - b.detach()).abs().max() / b.detach().abs().max()).item() <= tol

    assert close(y, y64, 1e-3)
    assert close(xd.grad, x64.grad, 2e-3) and close(ours.beta.grad, refd.beta.grad, 2e-3)
    assert close(ours.gamma.grad, refd.gamma.grad, 2e-3)
    # and the NCHW path on the same numbers gives the same result to TF32 accuracy
    xn = x.to(DEV).requires_grad_(True)
    yn = ours(xn)
    assert yn.is_contiguous() and torch.allclose(yn, y.contiguous(), rtol=2e-3, atol=2e-3)


def test_compressor_channels_last_matches_default_layout():
    """The -m 3 model with use_channels_last(): same loss and gradients as the NCHW run (convolutions pick other cuDNN
    kernels, so equality is numerical), likelihood shapes unchanged."""
    torch.manual_seed(44)
    tasks = ("rgb", "depth_euclidean", "normal")
    a = mm.build_compressor(3, tasks, 24, 40, lmbda=1e-2).to(DEV).train()
    b = mm.build_compressor(3, tasks, 24, 40, lmbda=1e-2)
    b.load_state_dict(a.state_dict())
    b.to(DEV).train().use_channels_last()
    batch = mm.synthetic_batch(tasks, 2, size=256, seed=21, device=DEV)
    nz = torch.rand(2, 120, 1, 1, device=DEV) - 0.5
    ny = torch.rand(2, 24, 1, 1, device=DEV) - 0.5
    losses = []
    for m in (a, b):
        c = m.model["compressor"]
        ef, gf = c.entropy_bottleneck.forward, c.gaussian_conditional.forward
        c.entropy_bottleneck.forward = lambda x, training=None, ef=ef: ef(x, training, noise=nz)
        c.gaussian_conditional.forward = lambda y, s, means=None, training=None, gf=gf: gf(y, s, means, training, noise=ny)
        with torch.backends.cudnn.flags(allow_tf32=False):
            x_hats, lik = m(batch)
            loss, _ = m.rate_distortion_loss(m._as_model_format(batch), x_hats, lik, "train")
            loss.backward()
        losses.append(loss.item())
        assert lik["y"].shape == (2, 24, 4, 4)
    assert abs(losses[0] - losses[1]) <= 2e-3 * abs(losses[0])
    pb = dict(b.named_parameters())
    worst = 0.0
    for n, p in a.named_parameters():
        if p.grad is None or n.endswith("quantiles"):
            continue
        d = p.grad.abs().max().item()
        if d > 1e-12:
            worst = max(worst, (p.grad - pb[n].grad.reshape(p.grad.shape)).abs().max().item() / d)
    assert worst < 3e-2, worst


def test_conv_bias_gradient_kernel_matches_torch():
    """(f1) The convolutions' bias gradient through mmnc_channel_sum: same numbers as torch's own reduction, forward
    output identical to nn.Conv2d / nn.ConvTranspose2d, bit-reproducible run to run."""
    torch.manual_seed(51)
    for shape in ((5, 7, 33, 20), (64, 50, 64, 64), (3, 300, 1, 1), (2, 1, 256, 256)):
        g = torch.randn(*shape, device=DEV)
        got, want = mm.ops.channel_sum(g), g.sum(dim=(0, 2, 3))
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-3 * g[:, 0].numel() ** 0.5)
        assert torch.equal(got, mm.ops.channel_sum(g))
    for ours, ref, x in ((mm.conv(6, 10), torch.nn.Conv2d(6, 10, 5, 2, 2), torch.randn(3, 6, 32, 32)),
                         (mm.deconv(6, 10), torch.nn.ConvTranspose2d(6, 10, 5, 2, 2, 1), torch.randn(3, 6, 16, 16))):
        ref.load_state_dict(ours.state_dict())
        ours.to(DEV), ref.to(DEV)
        xa, xb = x.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
        with torch.backends.cudnn.flags(allow_tf32=False):
            ya, yb = ours(xa), ref(xb)
            w = torch.randn_like(ya)
            ya.backward(w), yb.backward(w)
        assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6) and torch.allclose(xa.grad, xb.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(ours.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(ours.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4)
