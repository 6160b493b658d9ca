"""ctypes binding of libmmnc_b200.so (the C ABI declared in include/mmnc_b200.h).

There is no CPU fallback: if the library has not been built, or a tensor is not on a CUDA device, the ops raise.
`build()` compiles csrc/*.cu for sm_100a with nvcc (works without a GPU); the built .so lives in-tree next to this
file so that it travels with the repository snapshot.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libmmnc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mmnc_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--split-compile", "0",
]

_lock = threading.Lock()
_lib = None

c_f32p = ctypes.c_void_p  # device pointers travel as integers
I64, I32, U64, F32, VP = ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); one entry per symbol declared in include/mmnc_b200.h
SIGNATURES = {
    "mmnc_version": (I32, []),
    "mmnc_last_error": (ctypes.c_char_p, []),
    "mmnc_launch_count": (U64, []),
    "mmnc_device_sm_count": (I32, []),
    "mmnc_quantize_noise": (I32, [VP, I64, I32, VP, U64, U64, VP, VP]),
    "mmnc_quantize_dequantize": (I32, [VP, I64, I64, I64, VP, I32, VP, VP]),
    "mmnc_quantize_symbols": (I32, [VP, I64, I64, I64, VP, I32, VP, VP]),
    "mmnc_dequantize_symbols": (I32, [VP, I64, I64, I64, VP, I32, VP, VP]),
    "mmnc_eb_forward": (I32, [VP, I64, I64, I64, VP, VP, I32, VP, U64, U64, F32, I32, VP, VP, VP, VP]),
    "mmnc_eb_backward": (I32, [VP, I64, I64, I64, VP, VP, VP, VP, F32, I32, VP, VP, VP]),
    "mmnc_eb_logits": (I32, [VP, I64, I64, VP, VP, VP]),
    "mmnc_eb_aux_loss": (I32, [VP, I64, VP, VP, VP, VP, VP]),
    "mmnc_gc_forward": (I32, [VP, VP, VP, I64, I64, I64, I64, I32, VP, U64, U64, F32, F32, VP, VP, VP, VP]),
    "mmnc_gc_backward": (I32, [VP, VP, VP, I64, I64, I64, I64, VP, VP, VP, F32, F32, VP, VP, VP]),
    "mmnc_lnsum_forward": (I32, [VP, I64, I64, I64, VP, VP]),
    "mmnc_lnsum_backward": (I32, [VP, I64, I64, I64, VP, VP, VP]),
    "mmnc_distortion_forward": (I32, [VP, VP, I64, I32, F32, VP, VP]),
    "mmnc_distortion_backward": (I32, [VP, VP, I64, I32, F32, VP, VP, VP]),
    "mmnc_rd_epilogue": (I32, [VP, I64, VP, I64, VP, I32, VP, VP, F32, F32, VP, I32, VP, F32, VP, VP, VP, VP, VP, VP]),
    "mmnc_gdn_forward": (I32, [VP, I64, I64, I64, VP, VP, I32, I32, VP, VP]),
    "mmnc_gdn_backward_workspace_bytes": (ctypes.c_size_t, [I64, I64, I64, I32]),
    "mmnc_gdn_backward": (I32, [VP, VP, I64, I64, I64, VP, VP, I32, I32, VP, VP, VP, VP, ctypes.c_size_t, VP]),
    "mmnc_gdn_backward_variant": (I32, [VP, VP, I64, I64, I64, I32]),
    "mmnc_gdn_forward_variant": (I32, [VP, VP, I64, I64, I64, I32]),
    "mmnc_gdn_forward_raw": (I32, [VP, I64, I64, I64, VP, VP, F32, F32, F32, I32, I32, VP, VP]),
    "mmnc_ssim_workspace_floats": (ctypes.c_size_t, [I64, I32, I32]),
    "mmnc_ms_ssim_workspace_floats": (ctypes.c_size_t, [I64, I32, I32]),
    "mmnc_ms_ssim": (I32, [VP, VP, I64, I32, I32, F32, F32, F32, F32, VP, VP, VP]),
    "mmnc_ssim_scale": (I32, [VP, VP, I64, I32, I32, F32, F32, F32, F32, VP, VP, VP, VP, VP, VP]),
    "mmnc_gdn_forward_workspace_bytes": (ctypes.c_size_t, [I64, I64, I64, I32]),
    "mmnc_gdn_forward_ws": (I32, [VP, I64, I64, I64, VP, VP, I32, I32, VP, VP, ctypes.c_size_t, VP]),
    "mmnc_gdn_forward_raw_ws": (I32, [VP, I64, I64, I64, VP, VP, F32, F32, F32, I32, I32, VP, VP, ctypes.c_size_t, VP]),
    "mmnc_gdn_backward_raw": (I32, [VP, VP, I64, I64, I64, VP, VP, F32, F32, F32, I32, I32, VP, VP, VP, VP,
                                    ctypes.c_size_t, VP]),
    "mmnc_gdn_nhwc_supported": (I32, [I64, I64, I64, I32, I32]),
    "mmnc_gdn_forward_raw_nhwc": (I32, [VP, I64, I64, I64, VP, VP, F32, F32, F32, I32, I32, VP, VP]),
    "mmnc_gdn_backward_raw_nhwc": (I32, [VP, VP, I64, I64, I64, VP, VP, F32, F32, F32, I32, I32, VP, VP, VP, VP,
                                         ctypes.c_size_t, VP]),
    "mmnc_nonneg_reparam_forward": (I32, [VP, I64, F32, F32, VP, VP]),
    "mmnc_nonneg_reparam_backward": (I32, [VP, VP, I64, F32, VP, VP]),
    "mmnc_channel_sum_workspace_floats": (I64, [I64, I64, I64]),
    "mmnc_channel_sum": (I32, [VP, I64, I64, I64, VP, VP, VP]),
    "mmnc_bias_add": (I32, [VP, VP, I64, I64, I64, VP]),
    "mmnc_argmax_sse": (I32, [VP, VP, I64, I32, I64, VP, VP, VP]),
    "mmnc_prep_u8_hwc_to_f32_chw": (I32, [VP, I64, I64, I32, I32, F32, VP, VP]),
    "mmnc_prep_u16_to_f32": (I32, [VP, I64, F32, VP, VP]),
    "mmnc_prep_labels": (I32, [VP, I64, I32, I32, VP, VP, VP]),
    "mmnc_pmf_to_quantized_cdf_h": (I32, [ctypes.POINTER(ctypes.c_float), I32, I32, ctypes.POINTER(ctypes.c_uint32)]),
    "mmnc_build_indexes": (I32, [VP, I64, VP, I32, F32, VP, VP]),
    "mmnc_rans_slab_words": (I64, [I64]),
    "mmnc_rans_pack_tables": (I32, [VP, VP, I32, I32, VP, VP, I64, VP]),
    "mmnc_rans_encode_batch": (I32, [VP, VP, I64, I64, I64, VP, I64, VP, VP, VP, I32, VP, VP, I64, VP, VP]),
    "mmnc_rans_compact": (I32, [VP, I64, VP, I64, VP, VP, I64, VP]),
    "mmnc_rans_decode_batch": (I32, [VP, VP, VP, VP, I64, I64, I64, VP, I64, VP, VP, VP, I32, VP, VP, VP]),
}


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cpp")))


def is_stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + [HEADER_PATH]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libmmnc_b200.so (in-tree)."""
    if not force and not is_stale():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO_PATH] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise RuntimeError(
                        f"mmnc_b200: {SO_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (needs nvcc). There is no CPU or PyTorch fallback for the rate path.")
                handle = ctypes.CDLL(SO_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)  # AttributeError here = header and library disagree
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


class MmncError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().mmnc_last_error().decode("utf-8", "replace")
        kind = {-1: "invalid argument", -2: "CUDA error", -3: "unsupported"}.get(rc, f"error {rc}")
        raise MmncError(f"mmnc_b200 {kind}: {msg}")


def launch_count() -> int:
    return int(lib().mmnc_launch_count())
