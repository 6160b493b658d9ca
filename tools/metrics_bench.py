"""Time the step metrics (mtc.py:359-384) at the C4 validation shapes: MS-SSIM per task tensor, PSNR from the distortion
sums, the semantic argmax pass.  Usage: python tools/metrics_bench.py [batch]"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mmnc_b200 as mm
from mmnc_b200 import metrics


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = "cuda:0"
    torch.manual_seed(0)
    for C in (3, 1):
        x = torch.rand(B, C, 256, 256, device=dev) * 255
        y = (x + torch.randn_like(x) * 8).clamp(0, 255)
        ms = timeit(lambda: metrics.ms_ssim(x, y, data_range=255.0))
        ms_t = timeit(lambda: metrics.ms_ssim_torch(x, y, data_range=255.0))
        n = x.numel()
        print(f"ms_ssim ({B},{C},256,256): mmnc_ssim_scale {ms:8.3f} ms ({n * 8 * 4 / 3 / ms / 1e6:7.1f} GB/s of the 8 B/element x 5 scales "
              f"it has to read)   stock torch ops {ms_t:8.3f} ms")


if __name__ == "__main__":
    main()
