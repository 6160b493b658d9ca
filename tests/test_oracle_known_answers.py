"""Pins the oracle: self-derived known answers and the committed fixtures.
The reference has no golden vectors for this path and CompressAI is absent, so this is what there is to pin
(parity unpinned in the sense of DESIGN.md)."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from helpers import load_golden
from oracle import compressai_ref as R
from oracle import native


@pytest.fixture(scope="module")
def gc():
    m = R.GaussianConditional(None)
    m.update_scale_table(R.get_scale_table())
    return m


def test_known_answers_appendix_c(gc, golden_dir):
    known = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    assert list(gc._quantized_cdf.shape) == [64, 3133] == known["gc_table_shape"]
    assert int(gc._cdf_length.sum()) == 27256 == known["gc_ragged_entries"]
    assert (-gc._offset[:6]).tolist() == [1, 1, 1, 1, 2, 2]
    assert (-gc._offset[-3:]).tolist() == [1223, 1383, 1565]
    assert int(gc._cdf_length.max()) - 2 == 3131
    assert abs(-R.GaussianConditional._standardized_quantile(1e-9 / 2) - 6.1094102) < 1e-6
    eb = R.EntropyBottleneck(4)
    eb.update()
    assert eb._cdf_length.tolist() == [23] * 4 and eb._offset.tolist() == [-10] * 4
    assert abs(float(eb.target[2]) - 21.416413) < 1e-5
    assert abs(float(eb._matrix0[0, 0, 0]) - (-1.452127)) < 1e-5
    assert abs(float(eb._matrix4[0, 0, 0]) - (-0.128505)) < 1e-5
    t = gc._quantized_cdf.numpy().astype(np.uint64).ravel()
    crc = int(np.bitwise_xor.reduce(t * (np.arange(t.size, dtype=np.uint64) | 1)))
    assert crc == known["gc_table_crc"]


def test_cdf_rows_are_valid(gc):
    cdf, ln = gc._quantized_cdf.numpy(), gc._cdf_length.numpy()
    for r in range(64):
        row = cdf[r, : ln[r]]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) > 0).all()
        assert (cdf[r, ln[r]:] == 0).all()


def test_pmf_to_cdf_golden():
    fx = load_golden()
    for p, want in zip(fx["pmf"], fx["pmf_cdf"]):
        got = np.array(native.pmf_to_quantized_cdf(p.tolist(), 16))
        assert (got == want).all()
        assert got[0] == 0 and got[-1] == 65536 and (np.diff(got) > 0).all()


def test_rans_golden_streams(gc):
    fx = load_golden()
    cdf, ln, off = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    for sym, idx, want in zip(fx["rans_sym"], fx["rans_idx"], fx["rans_bytes"]):
        b = native.encode_with_indexes_np(sym, idx, cdf, ln, off)
        assert b == want.tobytes()
        assert (native.decode_with_indexes_np(b, idx, cdf, ln, off) == sym).all()


def test_rans_format_facts(gc):
    """A.7: state starts at 2^31; flush writes (lo, hi); a stream of always-certain... smallest stream is 8 bytes."""
    cdf, ln, off = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    b = native.encode_with_indexes_np([0, 0], [0, 0], cdf, ln, off)
    assert len(b) % 4 == 0 and len(b) >= 8
    lo, hi = np.frombuffer(b[:8], dtype=np.uint32)
    assert (int(hi) << 32 | int(lo)) >= 2 ** 31


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(-300000, 300000), st.integers(0, 63)), min_size=2, max_size=200))
def test_rans_roundtrip_property(pairs):
    gcm = test_rans_roundtrip_property.gc
    cdf, ln, off = gcm._quantized_cdf.numpy(), gcm._cdf_length.numpy(), gcm._offset.numpy()
    sym = np.array([p[0] for p in pairs], dtype=np.int32)
    idx = np.array([p[1] for p in pairs], dtype=np.int32)
    b = native.encode_with_indexes_np(sym, idx, cdf, ln, off)
    assert (native.decode_with_indexes_np(b, idx, cdf, ln, off) == sym).all()


_g = R.GaussianConditional(None)
_g.update_scale_table(R.get_scale_table())
test_rans_roundtrip_property.gc = _g


def test_faithful_and_lean_marshalling_agree(gc):
    torch.manual_seed(3)
    scales = torch.exp(torch.empty(3, 5, 4, 4).uniform_(-3, 4))
    y = torch.randn(3, 5, 4, 4) * scales
    idx = gc.build_indexes(scales)
    gc.marshalling = "faithful"
    a = gc.compress(y, idx)
    gc.marshalling = "lean"
    b = gc.compress(y, idx)
    gc.marshalling = "faithful"
    assert a == b
    assert torch.equal(gc.decompress(a, idx), torch.round(y))


def test_eb_golden_forward_and_strings():
    fx = load_golden()
    eb = R.EntropyBottleneck(6)
    eb.load_state_dict({k: torch.from_numpy(v) for k, v in fx["eb_state"].item().items()})
    z = torch.from_numpy(fx["eb_z"])
    eb.eval()
    out, lik = eb(z)
    assert np.array_equal(out.detach().numpy(), fx["eb_eval_out"])
    assert np.allclose(lik.detach().numpy(), fx["eb_eval_lik"], rtol=1e-6, atol=0)
    eb.train()
    out, lik = eb(z, noise=torch.from_numpy(fx["eb_noise"]))
    assert np.allclose(lik.detach().numpy(), fx["eb_train_lik"], rtol=1e-6, atol=0)
    s = eb.compress(z)
    assert [x for x in s] == [w.tobytes() for w in fx["eb_strings"]]
    assert torch.equal(eb.decompress(s, z.shape[-2:]), out.new_tensor(fx["eb_eval_out"]))


def test_eb_sign_vs_plain_switch():
    """A.3: the two forms agree near the mode and diverge in the upper tail (plain cancels to 0)."""
    lo, up = torch.tensor([17.0]), torch.tensor([19.0])
    s = R.EntropyBottleneck(1, likelihood_form="sign")._likelihood_from_logits(lo, up)
    p = R.EntropyBottleneck(1, likelihood_form="plain")._likelihood_from_logits(lo, up)
    assert p.item() == 0.0 and abs(s.item() - 3.58e-8) < 1e-9


def test_gc_golden():
    fx = load_golden()
    gcm = R.GaussianConditional(None)
    gcm.update_scale_table(R.get_scale_table())
    gcm.eval()
    yh, yl = gcm(torch.from_numpy(fx["gc_y_bcast"]), torch.from_numpy(fx["gc_scales"]))
    assert yl.shape == (3, 8, 4, 4) and yh.shape == (3, 8, 1, 1)  # the reference's broadcast
    assert np.allclose(yl.numpy(), fx["gc_lik_bcast"], rtol=1e-6)
    idx = gcm.build_indexes(torch.from_numpy(fx["gc_scales"]))
    assert np.array_equal(idx.numpy(), fx["gc_idx"])
    s = gcm.compress(torch.from_numpy(fx["gc_y"]), idx)
    assert s == [w.tobytes() for w in fx["gc_strings"]]


def test_compress_rejects_shape_mismatch(gc):
    """A.6 / Appendix B1: compress raises when y and indexes differ in shape (what happens at 256^2 at HEAD)."""
    with pytest.raises(ValueError):
        gc.compress(torch.zeros(2, 4, 1, 1), torch.zeros(2, 4, 4, 4, dtype=torch.int32))
