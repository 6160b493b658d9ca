"""Shared helpers for the test-suite (tests may import oracle/; the product never does)."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "rate_path_fixtures.npz"), allow_pickle=True)


def hostcheck():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    lib.hc_rans_encode.restype = ctypes.c_int64
    return lib


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pack_eb(eb) -> torch.Tensor:
    """(C, 58) raw parameters in the kernel's order, from any module with CompressAI's parameter names."""
    C, parts = eb.channels, []
    for i in range(5):
        parts.append(getattr(eb, f"_matrix{i}").detach().reshape(C, -1))
        parts.append(getattr(eb, f"_bias{i}").detach().reshape(C, -1))
        if i < 4:
            parts.append(getattr(eb, f"_factor{i}").detach().reshape(C, -1))
    return torch.cat(parts, dim=1).contiguous()


def perturb_eb_(eb, seed=7):
    """Makes the tanh gates live: matrices += N(0,0.3), factors ~ U[-1,1], sorted quantiles."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for i in range(5):
            m = getattr(eb, f"_matrix{i}")
            m.add_(torch.randn(m.shape, generator=g) * 0.3)
            b = getattr(eb, f"_bias{i}")
            b.copy_(torch.rand(b.shape, generator=g) - 0.5)
            if i < 4:
                f = getattr(eb, f"_factor{i}")
                f.copy_(torch.rand(f.shape, generator=g) * 2 - 1)
        eb.quantiles.copy_(torch.sort(torch.randn(eb.channels, 1, 3, generator=g) * 5, dim=2).values)
    return eb


def assert_likelihood_close(got, want, rtol=1e-5, floor_atol=2e-9, what="", want64=None, tail_rtol=5e-5,
                            rtol32=2e-5):
    """Likelihood parity (north_star: 1e-5 relative in fp32; SURVEY.md A.8 / 8d: floor-aware).

      bulk  (lik > 1e-3)        : relative error <= rtol (1e-5)
      tails (1e-6 < lik <= 1e-3): relative error <= tail_rtol (5e-5)
      floor (lik <= 1e-6)       : absolute error <= 2e-9

    Why tiers: the likelihood is a difference of two nearly equal sigmoids / erfcs, so the fp32 rounding of the
    arguments is amplified by |argument| / lik-width.  Measured on this oracle: the fp32 torch-CPU oracle itself
    sits up to 4.4e-6 (bulk) and 5.7e-6 (tails) from the float64 value of the same formula, and two independent
    fp32 evaluations differ from each other by up to ~1.3e-5.  When the float64 oracle (`want64`) is available the
    bars above are applied against it (the exact value), and the fp32 oracle is additionally held to `rtol32`
    (2e-5, the sum of both roundings) on the bulk."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    ref = want if want64 is None else want64.detach().double().cpu()
    rel = (got - ref).abs() / ref.clamp_min(1e-30)
    bulk, tail = ref > 1e-3, (ref > 1e-6) & (ref <= 1e-3)
    if bulk.any():
        assert rel[bulk].max().item() <= rtol, f"{what}: bulk max rel err {rel[bulk].max().item():.3e} > {rtol}"
    if tail.any():
        assert rel[tail].max().item() <= tail_rtol, f"{what}: tail max rel err {rel[tail].max().item():.3e}"
    if want64 is not None and bulk.any():
        rel32 = ((got - want).abs() / want.clamp_min(1e-30))[bulk]
        assert rel32.max().item() <= rtol32, f"{what}: max rel err vs fp32 oracle {rel32.max().item():.3e}"
    small = ref <= 1e-6
    if small.any():
        err = (got - want).abs()[small].max().item()
        assert err <= floor_atol, f"{what}: abs err {err:.3e} below 1e-6 exceeds {floor_atol}"


def eb_double(eb):
    """float64 copy of an oracle EntropyBottleneck (the exact value of the same formula)."""
    from oracle import compressai_ref as R

    d = R.EntropyBottleneck(eb.channels, likelihood_form=eb.likelihood_form).double()
    d.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in eb.state_dict().items()})
    return d
