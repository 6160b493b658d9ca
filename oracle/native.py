"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

ctypes view of oracle/liboracle_ref.so (plain-C restatement of CompressAI 1.2.4's two pybind11 modules,
`compressai._CXX.pmf_to_quantized_cdf` and `compressai.ans.RansEncoder/RansDecoder`).

The Python signatures mirror the pybind11 originals that the reference reaches through
/root/reference/src/models/multi_task_compressor.py:509,543,546 (Python lists in, bytes / lists out),
plus "lean" ndarray variants used by bench.py's CPU baseline (variant B).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_ref.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (called by __graft_entry__.build and lazily by load())."""
    src = os.path.join(_HERE, "rans_cdf_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.orc_pmf_to_quantized_cdf.restype = ctypes.c_int
        lib.orc_pmf_to_quantized_cdf.argtypes = [
            ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        lib.orc_rans_encode_with_indexes.restype = ctypes.c_int64
        lib.orc_rans_encode_with_indexes.argtypes = [
            i32p, i32p, ctypes.c_int64, i32p, ctypes.c_int, ctypes.c_int, i32p, i32p,
            ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8))]
        lib.orc_rans_decode_with_indexes.restype = ctypes.c_int
        lib.orc_rans_decode_with_indexes.argtypes = [
            ctypes.c_char_p, ctypes.c_int64, i32p, ctypes.c_int64, i32p, ctypes.c_int, ctypes.c_int,
            i32p, i32p, i32p]
        lib.orc_free.restype = None
        lib.orc_free.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = 16) -> List[int]:
    """`compressai._CXX.pmf_to_quantized_cdf(pmf: List[float], precision) -> List[int]`."""
    lib = load()
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib.orc_pmf_to_quantized_cdf(
        p.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(p.size), int(precision),
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise ValueError(f"pmf_to_quantized_cdf failed (code {rc})")
    return out.astype(np.int64).tolist()


def _dense(cdfs) -> np.ndarray:
    if isinstance(cdfs, np.ndarray):
        return _i32(cdfs)
    width = max(len(r) for r in cdfs)
    t = np.zeros((len(cdfs), width), dtype=np.int32)
    for i, r in enumerate(cdfs):
        t[i, : len(r)] = r
    return t


def encode_with_indexes_np(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
    """Lean variant: contiguous int32 arrays in, bytes out (one stream)."""
    lib = load()
    s, ix = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    t, sz, off = _dense(cdfs), _i32(cdfs_sizes).reshape(-1), _i32(offsets).reshape(-1)
    if s.size != ix.size:
        raise ValueError("symbols / indexes size mismatch")
    outp = ctypes.POINTER(ctypes.c_uint8)()
    n = lib.orc_rans_encode_with_indexes(_p(s), _p(ix), int(s.size), _p(t), int(t.shape[0]), int(t.shape[1]),
                                         _p(sz), _p(off), ctypes.byref(outp))
    if n < 0:
        raise ValueError(f"rans encode failed (code {n})")
    data = ctypes.string_at(outp, n)
    lib.orc_free(outp)
    return data


def decode_with_indexes_np(encoded: bytes, indexes, cdfs, cdfs_sizes, offsets) -> np.ndarray:
    lib = load()
    ix = _i32(indexes).reshape(-1)
    t, sz, off = _dense(cdfs), _i32(cdfs_sizes).reshape(-1), _i32(offsets).reshape(-1)
    out = np.empty(ix.size, dtype=np.int32)
    rc = lib.orc_rans_decode_with_indexes(encoded, len(encoded), _p(ix), int(ix.size), _p(t), int(t.shape[0]),
                                          int(t.shape[1]), _p(sz), _p(off), _p(out))
    if rc != 0:
        raise ValueError(f"rans decode failed (code {rc})")
    return out


class RansEncoder:
    """`compressai.ans.RansEncoder` look-alike: Python lists in (as CompressAI marshals them), bytes out."""

    def encode_with_indexes(self, symbols: List[int], indexes: List[int], cdfs: List[List[int]],
                            cdfs_sizes: List[int], offsets: List[int]) -> bytes:
        return encode_with_indexes_np(symbols, indexes, cdfs, cdfs_sizes, offsets)


class RansDecoder:
    """`compressai.ans.RansDecoder` look-alike."""

    def decode_with_indexes(self, encoded: bytes, indexes: List[int], cdfs: List[List[int]],
                            cdfs_sizes: List[int], offsets: List[int]) -> List[int]:
        return decode_with_indexes_np(encoded, indexes, cdfs, cdfs_sizes, offsets).tolist()
