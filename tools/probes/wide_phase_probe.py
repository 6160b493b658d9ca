"""Per-phase clock totals of the wide-layer GDN kernels (gdn_tc_wide.cu built with -DMMNC_WIDE_PROFILE).

Build here (no GPU needed), run on the GPU box:
    python tools/probes/wide_phase_probe.py --build
    gpurun -- python tools/probes/wide_phase_probe.py 64 256 64 64
CTA 0's leader thread accumulates clock64() deltas between the phase boundaries of every tile it processes; the
numbers are cycles per tile (forward: 0 A fill incl. waiting for x, 1 barrier, 2 MMA issue, 3 MMA wait, 4 epilogue,
5 end barrier; dx kernel: 8 A fill, 9 barrier, 10 MMA1 issue, 11 g request + MMA1 wait, 12 epilogue 1, 13 barrier,
14 MMA2 issue, 15 MMA2 wait, 16 epilogue 2)."""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "multi-modal-neural-compression_b200", "csrc")
SO = os.path.join(ROOT, "tools", "probes", "libmmnc_prof.so")


def build():
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "--split-compile", "0", "-DMMNC_WIDE_PROFILE", "-o", SO] + src
    subprocess.run(cmd, check=True)


def main():
    if "--build" in sys.argv:
        build()
        return
    import torch
    B, C, H, W = (int(v) for v in sys.argv[1:5])
    L = ctypes.CDLL(SO)
    L.mmnc_gdn_forward_workspace_bytes.restype = ctypes.c_size_t
    L.mmnc_gdn_forward_workspace_bytes.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_int]
    L.mmnc_gdn_backward_workspace_bytes.restype = ctypes.c_size_t
    L.mmnc_gdn_backward_workspace_bytes.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_int]
    VP, I64, I32, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    L.mmnc_gdn_forward_ws.argtypes = [VP, I64, I64, I64, VP, VP, I32, I32, VP, VP, SZ, VP]
    L.mmnc_gdn_backward.argtypes = [VP, VP, I64, I64, I64, VP, VP, I32, I32, VP, VP, VP, VP, SZ, VP]
    L.mmnc_last_error.restype = ctypes.c_char_p
    dev = "cuda:0"
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device=dev)
    g = torch.randn(B, C, H, W, device=dev)
    beta = torch.rand(C, device=dev) + 0.5
    gamma = torch.rand(C, C, device=dev) * 0.05
    y, dx = torch.empty_like(x), torch.empty_like(x)
    db, dg = torch.empty_like(beta), torch.empty_like(gamma)
    HW = H * W
    nf = L.mmnc_gdn_forward_workspace_bytes(B, C, HW, 3)
    nb = L.mmnc_gdn_backward_workspace_bytes(B, C, HW, 3)
    wf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev)
    wb = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = (ctypes.c_ulonglong * 32)()
    reps = 5
    tiles = (B * HW + 127) // 128
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    per_cta0 = (tiles + sms - 1) // sms  # CTA 0 always takes the ceiling
    for it in range(2):
        L.mmnc_wide_profile_read(out)
        for _ in range(reps):
            rc = L.mmnc_gdn_forward_ws(x.data_ptr(), B, C, HW, beta.data_ptr(), gamma.data_ptr(), 0, 3, y.data_ptr(),
                                       wf.data_ptr(), nf, None)
            assert rc == 0, L.mmnc_last_error()
            rc = L.mmnc_gdn_backward(x.data_ptr(), g.data_ptr(), B, C, HW, beta.data_ptr(), gamma.data_ptr(), 0, 3,
                                     dx.data_ptr(), db.data_ptr(), dg.data_ptr(), wb.data_ptr(), nb, None)
            assert rc == 0, L.mmnc_last_error()
        torch.cuda.synchronize()
        L.mmnc_wide_profile_read(out)
        if it == 1:
            n = reps * per_cta0
            fwd = [out[i] / n for i in range(6)]
            bwd = [out[i] / n for i in range(8, 17)]
            print("tiles per CTA", per_cta0)
            print("forward cycles/tile:", [round(v) for v in fwd], "sum", round(sum(fwd)))
            print("leader, cycles/tile over the forward + both dx contractions: waiting for chunks", round(out[20] / n),
                  " requesting chunks", round(out[22] / n), " issuing MMAs + commits", round(out[23] / n))
            print("dx kernel, cycles per launch: slowest CTA (max over launches)", out[30], " mean CTA", round(out[31] / reps / min(sms, tiles)))
            print("dx      cycles/tile:", [round(v) for v in bwd], "sum", round(sum(bwd)))


if __name__ == "__main__":
    main()
