"""Kernel-level timing of the GDN contraction (forward / backward) for the layer shapes of BASELINE config C2 and C3.
usage: python tools/gdn_bench.py [--precisions fp32,tf32,3xtf32] [--backward] [--shapes B,C,H,W ...]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precisions", default="fp32,tf32,3xtf32")
ap.add_argument("--backward", action="store_true")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--shapes", nargs="*", default=["64,50,256,256", "64,100,128,128", "64,100,64,64", "64,50,64,64",
                                                "64,64,256,256", "64,128,128,128", "64,33,8,8", "64,3,256,256"])
args = ap.parse_args()
dev = torch.device("cuda:0")
peak = 6547.5
for shp in args.shapes:
    B, C, H, W = map(int, shp.split(","))
    x = [torch.randn(B, C, H, W, device=dev) for _ in range(2)]
    g = torch.randn(B, C, H, W, device=dev)
    beta = 1 + torch.rand(C, device=dev) * 0.5
    gamma = 0.1 * torch.eye(C, device=dev) + torch.rand(C, C, device=dev) * 0.01
    n = x[0].numel()
    for prec in args.precisions.split(","):
        def fwd(i=[0]):
            i[0] += 1
            return mm.ops.gdn(x[i[0] & 1], beta, gamma, False, prec)
        with torch.no_grad():
            fwd(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fwd()
            e1.record(); torch.cuda.synchronize()
        tf = e0.elapsed_time(e1) / args.reps * 1e-3
        line = f"{shp:>16} {prec:>7} fwd {tf*1e3:8.3f} ms {8*n/tf/1e9:7.0f} GB/s ({8*n/tf/1e9/peak:.3f})"
        if args.backward:
            xr, br, gr = x[0].requires_grad_(True), beta.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
            def fb():
                torch.autograd.grad(mm.ops.gdn(xr, br, gr, False, prec), [xr, br, gr], g)
            fb(); torch.cuda.synchronize()
            e0.record()
            for _ in range(max(2, args.reps // 2)):
                fb()
            e1.record(); torch.cuda.synchronize()
            tb = e0.elapsed_time(e1) / max(2, args.reps // 2) * 1e-3 - tf
            line += f" | bwd {tb*1e3:8.3f} ms {12*n/tb/1e9:7.0f} GB/s ({12*n/tb/1e9/peak:.3f})"
        print(line, flush=True)
    del x, g
