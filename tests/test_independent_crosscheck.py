"""Independent cross-checks of the parts where product and oracle share ancestry ("compares a file with
its twin").  CompressAI itself is not installable here, so nothing below can pin parity to its binaries; what these
tests do is take the oracle's authorship out of the loop:

  * everything is re-derived in numpy / scipy (scipy.special.erfc, scipy.special.expit, scipy.stats.norm) and plain
    Python integers, written from the published algorithms (CompressAI ops.cpp / entropy_models.py / rans_interface.cpp
    and ryg_rans rans64.h), without importing `oracle/` for the value under test;
  * the product's Gaussian CDF tables (`update()` + the native `pmf_to_quantized_cdf`) must come out BIT-EXACT from the
    fp32 numpy / scipy evaluation in CompressAI's op order, and within 2 counts of 65536 from a float64 evaluation;
    the bottleneck tables (tanh / softplus network) within 2 counts in both precisions, lengths and offsets exactly;
  * the oracle's likelihood formulas must agree with scipy's float64 special functions to 1e-10;
  * the C coder of the oracle must produce the bytes of a pure-Python big-integer rANS.

No GPU needed.
"""
import numpy as np
import pytest
import scipy.special
import scipy.stats
import torch

import mmnc_b200 as mm
from helpers import perturb_eb_


# ------------------------------------------------------------------------------------------------ pmf -> CDF
def pmf_to_quantized_cdf_np(pmf, precision=16):
    """CompressAI ops.cpp `pmf_to_quantized_cdf`, re-written with numpy integers."""
    pmf = np.asarray(pmf, dtype=np.float32)
    n = len(pmf)
    scaled = (pmf * np.float32(1 << precision)).astype(np.float32)
    rounded = np.floor(np.abs(scaled) + np.float32(0.5)).astype(np.int64) * np.sign(scaled).astype(np.int64)  # std::round
    cdf = np.zeros(n + 1, dtype=np.int64)
    cdf[1:] = rounded
    total = int(cdf.sum())
    if total == 0:
        raise ValueError("all-zero pmf")
    cdf = ((1 << precision) * cdf) // total
    cdf = np.cumsum(cdf)
    cdf[-1] = 1 << precision
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            freqs = cdf[1:] - cdf[:-1]
            cand = np.where(freqs > 1)[0]
            assert len(cand), "no frequency left to steal"
            j = cand[np.argmin(freqs[cand])]  # smallest frequency above 1, first one on ties
            if j < i:
                cdf[j + 1:i + 1] -= 1
            else:
                cdf[i + 1:j + 1] += 1
    return cdf


def test_native_pmf_to_quantized_cdf_matches_independent_numpy_version():
    rng = np.random.default_rng(3)
    cases = [np.array([0.5, 0.5]), np.array([1.0, 0.0, 0.0, 0.0]), np.array([1e-9] * 40 + [1.0]),
             np.array([0.25, 0.0, 0.5, 0.0, 0.25])]
    for _ in range(200):
        n = int(rng.integers(2, 300))
        p = rng.random(n) ** int(rng.integers(1, 12))
        p[rng.random(n) < 0.2] = 0.0  # zero bins force the frequency-stealing loop
        if p.sum() == 0:
            p[0] = 1.0
        cases.append(p / p.sum())
    for p in cases:
        p32 = p.astype(np.float32)
        want = pmf_to_quantized_cdf_np(p32)
        got = mm.ops.pmf_to_quantized_cdf(p32.tolist(), 16)
        assert got == want.tolist()
        assert all(b > a for a, b in zip(got, got[1:])) and got[0] == 0 and got[-1] == 65536


# ------------------------------------------------------------------------------------------------ Gaussian tables
def _gc_rows(dtype):
    table = mm.get_scale_table().numpy()
    mult = -scipy.stats.norm.ppf(1e-9 / 2)
    center = np.ceil(table.astype(np.float64) * mult).astype(np.int64)
    rows = []
    for i, s in enumerate(table.astype(dtype)):
        k = np.abs(np.arange(2 * center[i] + 1) - center[i]).astype(dtype)
        if dtype == np.float32:  # CompressAI's op order: 0.5 * erfc(-(2 ** -0.5) * x), everything in fp32
            c = np.float32(-(2 ** -0.5))
            up = np.float32(0.5) * scipy.special.erfc(c * ((np.float32(0.5) - k) / s)).astype(np.float32)
            lo = np.float32(0.5) * scipy.special.erfc(c * ((np.float32(-0.5) - k) / s)).astype(np.float32)
        else:
            up, lo = scipy.stats.norm.cdf((0.5 - k) / s), scipy.stats.norm.cdf((-0.5 - k) / s)
        pmf, tail = up - lo, 2 * lo[:1]
        rows.append(pmf_to_quantized_cdf_np(np.concatenate([pmf, tail]).astype(np.float32)))
    return center, rows


@pytest.fixture(scope="module")
def gc_tables():
    gc = mm.GaussianConditional(None)
    gc.update_scale_table(mm.get_scale_table())
    return gc


def test_gaussian_tables_bit_exact_from_fp32_numpy_scipy(gc_tables):
    center, rows = _gc_rows(np.float32)
    assert np.array_equal(-center, gc_tables._offset.numpy())
    assert np.array_equal(2 * center + 3, gc_tables._cdf_length.numpy())
    assert gc_tables._quantized_cdf.shape == (64, 3133) and sum(len(r) for r in rows) == 27256
    for i, row in enumerate(rows):
        assert np.array_equal(row, gc_tables._quantized_cdf[i, :len(row)].numpy()), f"row {i}"
        assert not gc_tables._quantized_cdf[i, len(row):].any()


def test_gaussian_tables_within_two_counts_of_float64_scipy(gc_tables):
    """The fp32 pmf CompressAI evaluates differs from the exact one by ~1e-7: a few hundred of the 27 256 entries move
    by one or two counts of 65536 (a code-length effect below 1e-4 bits per symbol)."""
    center, rows = _gc_rows(np.float64)
    diff = np.concatenate([np.abs(row - gc_tables._quantized_cdf[i, :len(row)].numpy()) for i, row in enumerate(rows)])
    assert diff.max() <= 2 and (diff > 0).mean() < 0.05


# ------------------------------------------------------------------------------------------------ bottleneck tables
def _eb_tables_np(sd, dtype):
    f = lambda k: sd[k].numpy().astype(dtype)  # noqa: E731
    q = f("quantiles")
    med = q[:, 0, 1]
    minima = np.maximum(np.ceil(med - q[:, 0, 0]).astype(np.int32), 0)
    maxima = np.maximum(np.ceil(q[:, 0, 2] - med).astype(np.int32), 0)
    pmf_start, pmf_length = med - minima.astype(dtype), maxima + minima + 1
    samples = np.arange(int(pmf_length.max())).astype(dtype)[None, :] + pmf_start[:, None]

    def logits(v):  # (C, L) -> (C, L): five layers of softplus(matrix) @ h + bias, gated by tanh(factor) * tanh(h)
        h = v[:, None, :]
        for i in range(5):
            m, b = f(f"_matrix{i}"), f(f"_bias{i}")
            sp = np.where(m > 20, m, np.log1p(np.exp(m))).astype(dtype)
            acc = None
            for j in range(sp.shape[2]):
                term = sp[:, :, j:j + 1] * h[:, j:j + 1, :]
                acc = term if acc is None else (acc + term).astype(dtype)
            h = (acc + b).astype(dtype)
            if i < 4:
                h = (h + np.tanh(f(f"_factor{i}")) * np.tanh(h)).astype(dtype)
        return h[:, 0, :]

    half = dtype(0.5)
    lo, up = logits((samples - half).astype(dtype)), logits((samples + half).astype(dtype))
    sign = -np.sign(lo + up)
    sig = scipy.special.expit
    pmf = np.abs(sig(sign * up) - sig(sign * lo)).astype(dtype)
    tail = (sig(lo[:, :1]) + sig(-up[:, -1:])).astype(dtype)
    rows = [pmf_to_quantized_cdf_np(np.concatenate([pmf[c, :pmf_length[c]], tail[c]]).astype(np.float32))
            for c in range(q.shape[0])]
    return -minima, pmf_length + 2, rows


@pytest.mark.parametrize("seed", [7, 8])
def test_bottleneck_tables_from_numpy_scipy(seed):
    """Offsets and lengths exactly; CDF entries within 2 counts of 65536 of a numpy evaluation in fp32 (numpy's libm
    tanh / exp differ from torch's vectorised ones in the last ulp, which moves a few roundings) and in float64."""
    eb = mm.EntropyBottleneck(96)
    perturb_eb_(eb, seed)
    eb.update()
    for dtype in (np.float32, np.float64):
        off, ln, rows = _eb_tables_np(eb.state_dict(), dtype)
        assert np.array_equal(off, eb._offset.numpy()) and np.array_equal(ln, eb._cdf_length.numpy())
        diff = np.concatenate([np.abs(r - eb._quantized_cdf[c, :len(r)].numpy()) for c, r in enumerate(rows)])
        assert diff.max() <= 2 and (diff > 0).mean() < 0.05, ((diff > 0).sum(), diff.size, diff.max())


def test_known_constants_from_scipy():
    """The self-derived known answers of SURVEY.md Appendix C, recomputed with scipy instead of the oracle."""
    assert abs(-scipy.stats.norm.ppf(1e-9 / 2) - 6.109410) < 1e-6
    assert abs(np.log(2 / 1e-9 - 1) - 21.416413) < 1e-6
    scale = 10 ** (1 / 5)
    assert abs(np.log(np.expm1(1 / scale / 3)) - (-1.452127)) < 1e-6
    assert abs(np.log(np.expm1(1 / scale / 1)) - (-0.128505)) < 1e-6
    eb = mm.EntropyBottleneck(4)
    assert abs(float(eb._matrix0[0, 0, 0]) + 1.452127) < 1e-5 and abs(float(eb._matrix4[0, 0, 0]) + 0.128505) < 1e-5
    assert abs(float(eb.target[2]) - 21.416413) < 1e-5


# ------------------------------------------------------------------------------------------------ likelihood formulas
def test_oracle_likelihood_formulas_against_scipy_float64():
    from oracle import compressai_ref as R

    rng = np.random.default_rng(11)
    # Gaussian conditional, eval mode (round, no means): lik = Phi((0.5 - |y|) / s) - Phi((-0.5 - |y|) / s), floored
    scales = np.exp(rng.uniform(np.log(0.05), np.log(64), 4000))
    y = rng.standard_normal(4000) * scales
    gc = R.GaussianConditional(None).double().eval()
    _, lik = gc(torch.from_numpy(y).reshape(1, -1, 1, 1), torch.from_numpy(scales).reshape(1, -1, 1, 1))
    s = np.maximum(scales, 0.11)
    v = np.abs(np.round(y))
    want = np.maximum(scipy.stats.norm.cdf((0.5 - v) / s) - scipy.stats.norm.cdf((-0.5 - v) / s), 1e-9)
    got = lik.detach().reshape(-1).numpy()
    big = want > 1e-7
    assert np.max(np.abs(got - want)[big] / want[big]) < 1e-9
    # entropy bottleneck: the MLP in numpy float64 + expit
    eb = R.EntropyBottleneck(8).double()
    perturb_eb_(eb, 5)
    eb.double().eval()
    z = rng.standard_normal((3, 8, 5)) * 4
    _, lik = eb(torch.from_numpy(z))
    sd = {k: v.detach() for k, v in eb.state_dict().items()}
    f = lambda k: sd[k].numpy().astype(np.float64)  # noqa: E731
    med = f("quantiles")[:, 0, 1]
    zq = np.round(z - med[None, :, None]) + med[None, :, None]

    def logits(v):  # v (B, C, L)
        h = v[:, :, None, :]
        for i in range(5):
            m, b = f(f"_matrix{i}"), f(f"_bias{i}")
            h = np.einsum("cij,bcjl->bcil", np.logaddexp(0, m), h) + b[None]
            if i < 4:
                h = h + np.tanh(f(f"_factor{i}"))[None] * np.tanh(h)
        return h[:, :, 0, :]

    lo, up = logits(zq - 0.5), logits(zq + 0.5)
    want = np.maximum(np.abs(scipy.special.expit(up) - scipy.special.expit(lo)), 1e-9)
    got = lik.detach().numpy()
    big = want > 1e-7
    assert np.max(np.abs(got - want)[big] / want[big]) < 1e-9


# ------------------------------------------------------------------------------------------------ rANS
def rans_encode_py(symbols, indexes, cdfs, cdf_sizes, offsets):
    """CompressAI rans_interface.cpp `encode_with_indexes` over ryg_rans rans64.h, in Python integers."""
    L, prec, bbits = 1 << 31, 16, 4
    syms = []  # (start, range, bypass)
    for s, ci in zip(symbols, indexes):
        cdf, max_value = cdfs[ci], int(cdf_sizes[ci]) - 2
        value, raw = int(s) - int(offsets[ci]), 0
        if value < 0:
            raw, value = -2 * value - 1, max_value
        elif value >= max_value:
            raw, value = 2 * (value - max_value), max_value
        syms.append((int(cdf[value]), int(cdf[value + 1]) - int(cdf[value]), False))
        if value == max_value:
            nb = 0
            while (raw >> (nb * bbits)) != 0:
                nb += 1
            val = nb
            while val >= 15:
                syms.append((15, 0, True))
                val -= 15
            syms.append((val, 0, True))
            for j in range(nb):
                syms.append(((raw >> (j * bbits)) & 15, 0, True))
    x, words = L, []
    for start, rng, bypass in reversed(syms):
        freq = (1 << (prec - bbits)) if bypass else rng
        if x >= ((L >> prec) << 32) * freq:
            words.append(x & 0xFFFFFFFF)
            x >>= 32
        x = ((x << bbits) | start) if bypass else ((x // freq) << prec) + (x % freq) + start
    words += [x >> 32, x & 0xFFFFFFFF]
    return b"".join(int(w).to_bytes(4, "little") for w in reversed(words))


def test_c_coder_matches_pure_python_rans(gc_tables):
    from oracle import native

    rng = np.random.default_rng(9)
    cdf = gc_tables._quantized_cdf.numpy()
    ln, off = gc_tables._cdf_length.numpy(), gc_tables._offset.numpy()
    table = mm.get_scale_table().numpy()
    for n in (1, 2, 33, 400):
        idx = rng.integers(0, 64, n).astype(np.int32)
        sym = np.round(rng.standard_normal(n) * table[idx] * 1.5).astype(np.int32)
        esc = rng.random(n) < 0.1
        sym[esc] = rng.integers(-2 ** 18, 2 ** 18, int(esc.sum()))
        want = rans_encode_py(sym, idx, cdf, ln, off)
        assert native.encode_with_indexes_np(sym, idx, cdf, ln, off) == want
