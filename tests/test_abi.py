"""The C-ABI library loads and exports every symbol include/mmnc_b200.h declares (no compute without a GPU)."""
import ctypes
import re

import pytest

import mmnc_b200 as mm


def _declared():
    text = open(mm._lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmnc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 28
    lib = ctypes.CDLL(mm._lib.SO_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but missing from the library"
    assert sorted(mm._lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_library_basics():
    lib = mm._lib.lib()
    assert lib.mmnc_version() >= 100
    assert isinstance(mm.launch_count(), int)
    assert lib.mmnc_rans_slab_words(300) >= (300 * 52) // 32 + 4


def test_argument_validation_without_gpu():
    lib = mm._lib.lib()
    rc = lib.mmnc_gc_forward(None, None, None, 2, 3, 5, 4, 0, None, 0, 0, 0.11, 1e-9, None, None, None, None)
    assert rc == -1 and b"spatial" in lib.mmnc_last_error()
    rc = lib.mmnc_eb_forward(None, 1, 1, 1, None, None, 9, None, 0, 0, 1e-9, 0, None, None, None, None)
    assert rc == -1
    assert lib.mmnc_rans_encode_batch(None, None, 0, 1, 1, None, 0, None, None, None, 0, None, None, 100, None, None) == -1
    assert lib.mmnc_rans_decode_batch(None, None, None, None, 0, 1, 1, None, 0, None, None, None, 0, None, None, None) == -1
    assert lib.mmnc_rans_pack_tables(None, None, 0, 0, None, None, 0, None) == -1
    # empty inputs are fine and launch nothing
    before = mm.launch_count()
    assert lib.mmnc_eb_forward(None, 0, 4, 1, None, None, 0, None, 0, 0, 1e-9, 0, None, None, None, None) == 0
    assert lib.mmnc_gdn_forward(None, 0, 8, 16, None, None, 0, 0, None, None) == 0
    assert mm.launch_count() == before


def test_pmf_to_quantized_cdf_host_function_matches_oracle():
    import numpy as np

    from helpers import load_golden
    from oracle import native

    fx = load_golden()
    for p, want in zip(fx["pmf"], fx["pmf_cdf"]):
        assert mm.ops.pmf_to_quantized_cdf(p.tolist(), 16) == want.tolist()
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(2, 400))
        p = rng.random(n).astype(np.float32) ** int(rng.integers(1, 10))
        p /= p.sum()
        assert mm.ops.pmf_to_quantized_cdf(p.tolist(), 16) == native.pmf_to_quantized_cdf(p.tolist(), 16)
    with pytest.raises(ValueError):
        mm.ops.pmf_to_quantized_cdf([0.5, float("nan")], 16)
    with pytest.raises(ValueError):
        mm.ops.pmf_to_quantized_cdf([0.0, 0.0], 16)


def test_no_cpu_fallback():
    import torch

    eb = mm.EntropyBottleneck(4)
    with pytest.raises(RuntimeError, match="no CPU"):
        eb(torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU"):
        mm.GDN(4)(torch.zeros(1, 4, 2, 2))
    gc = mm.GaussianConditional(None)
    with pytest.raises(RuntimeError, match="no CPU"):
        gc(torch.zeros(1, 4, 2, 2), torch.ones(1, 4, 2, 2))


def test_product_never_imports_oracle():
    import os
    import re

    root = os.path.dirname(mm._lib.SO_PATH)
    pat = re.compile(r"^\s*(from\s+\.*oracle|import\s+oracle|#\s*include\s+[\"<][^\n]*oracle)|"
                     r"import_module\([^)]*oracle|liboracle|CDLL\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} imports / links the oracle"
