"""Times the GDN backward C-ABI call alone (CUDA events on the launching stream), per kernel variant.
usage: python tools/gdn_bwd_bench.py [--shapes B,C,H,W ...] [--reps 10]      (MMNC_GDN_BWD=v1|v2 forces a variant)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--shapes", nargs="*", default=["64,50,256,256", "64,100,128,128", "64,100,64,64", "64,64,256,256",
                                                "64,100,32,32", "64,50,128,128"])
args = ap.parse_args()
dev = torch.device("cuda:0")
L = mm._lib.lib()
peak = 6547.5
for shp in args.shapes:
    B, C, H, W = map(int, shp.split(","))
    HW = H * W
    xs = [torch.randn(B, C, H, W, device=dev) for _ in range(2)]
    gs = [torch.randn(B, C, H, W, device=dev) for _ in range(2)]
    beta = 1 + torch.rand(C, device=dev) * 0.5
    gamma = 0.1 * torch.eye(C, device=dev) + torch.rand(C, C, device=dev) * 0.01
    dx = torch.empty_like(xs[0])
    db, dg = torch.empty_like(beta), torch.empty_like(gamma)
    nbytes = int(L.mmnc_gdn_backward_workspace_bytes(B, C, HW, 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    variant = L.mmnc_gdn_backward_variant(xs[0].data_ptr(), gs[0].data_ptr(), B, C, HW, 1)

    def run(i):
        mm._lib.check(L.mmnc_gdn_backward(xs[i & 1].data_ptr(), gs[i & 1].data_ptr(), B, C, HW, beta.data_ptr(),
                                          gamma.data_ptr(), 0, 1, dx.data_ptr(), db.data_ptr(), dg.data_ptr(),
                                          ws.data_ptr(), nbytes, st))
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / args.reps * 1e-3
    n = xs[0].numel()
    print(f"{shp:>16} variant {variant} bwd {t*1e3:8.3f} ms {12*n/t/1e9:7.0f} GB/s ({12*n/t/1e9/peak:.3f})", flush=True)
    del xs, gs, dx
