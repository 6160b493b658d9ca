"""Launch-bound layers: time the GDN forward / backward C-ABI calls on the small shapes of BASELINE config C2
(ctypes calls in a tight loop, CUDA events around 200 launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm  # noqa: E402

dev = torch.device("cuda:0")
L = mm._lib.lib()
shapes = [(64, 100, 8, 8), (64, 100, 16, 16), (64, 100, 32, 32), (64, 33, 2, 2), (64, 33, 4, 4), (64, 33, 8, 8),
          (64, 300, 4, 4), (64, 300, 2, 2), (64, 300, 1, 1), (64, 50, 32, 32), (64, 50, 64, 64), (64, 3, 128, 128)]
if len(sys.argv) > 1:
    shapes = [tuple(map(int, a.split(","))) for a in sys.argv[1:]]
st = torch.cuda.current_stream().cuda_stream
for B, C, H, W in shapes:
    HW = H * W
    x, g = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
    beta = 1 + torch.rand(C, device=dev) * 0.5
    gamma = 0.1 * torch.eye(C, device=dev) + torch.rand(C, C, device=dev) * 0.01
    y, dx = torch.empty_like(x), torch.empty_like(x)
    db, dg = torch.empty_like(beta), torch.empty_like(gamma)
    nbytes = int(L.mmnc_gdn_backward_workspace_bytes(B, C, HW, 3))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def fwd():
        mm._lib.check(L.mmnc_gdn_forward(x.data_ptr(), B, C, HW, beta.data_ptr(), gamma.data_ptr(), 0, 3, y.data_ptr(), st))

    def bwd():
        mm._lib.check(L.mmnc_gdn_backward(x.data_ptr(), g.data_ptr(), B, C, HW, beta.data_ptr(), gamma.data_ptr(), 0, 3,
                                          dx.data_ptr(), db.data_ptr(), dg.data_ptr(), ws.data_ptr(), nbytes, st))
    out = []
    for fn in (fwd, bwd):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 200 * 1e3)
    variant = L.mmnc_gdn_backward_variant(x.data_ptr(), g.data_ptr(), B, C, HW, 3)
    print(f"{B},{C},{H},{W}: fwd {out[0]:7.1f} us   bwd {out[1]:7.1f} us (variant {variant})", flush=True)
