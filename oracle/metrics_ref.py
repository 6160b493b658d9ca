"""TEST INFRASTRUCTURE (oracle): CPU restatement of the two metrics the reference logs on every step
(/root/reference/src/models/multi_task_compressor.py:92 `self.metrics = {"psnr": peak_signal_noise_ratio, "ms-ssim":
ms_ssim}`, :359-384 `average_metrics`).  Neither library is vendored in /root/reference nor installed here:

  * torchmetrics.functional.peak_signal_noise_ratio (requirements.txt: torchmetrics) - published definition
    10 log10(data_range^2 / mean((pred - target)^2)), the mean over every element of the batch;
  * pytorch_msssim.ms_ssim (requirements.txt: pytorch-msssim) - published algorithm: 11-tap Gaussian window, sigma 1.5,
    built in fp32 and normalised; 'valid' separable filtering as two grouped conv2d; K1, K2 = 0.01, 0.03; five scales
    with avg_pool2d(kernel 2, padding = side % 2) in between; relu of the contrast terms of scales 0-3 and of the SSIM of
    scale 4; weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333); mean over channels, then over the batch.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  Parity of this restatement is pinned
by tests/test_metrics.py against a float64 numpy / scipy.ndimage implementation written independently of it."""
from typing import Dict

import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def peak_signal_noise_ratio(pred: torch.Tensor, target: torch.Tensor, data_range: float) -> torch.Tensor:
    mse = torch.mean((pred - target) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(data_range, dtype=mse.dtype) ** 2 / mse)


def _window(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gaussian_filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    C = x.shape[1]
    x = F.conv2d(x, win.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
    return F.conv2d(x, win.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)


def _ssim(x, y, data_range, win):
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _gaussian_filter(x, win), _gaussian_filter(y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = _gaussian_filter(x * x, win) - mu1_sq
    sigma2_sq = _gaussian_filter(y * y, win) - mu2_sq
    sigma12 = _gaussian_filter(x * y, win) - mu1_mu2
    cs_map = (2 * sigma12 + c2) / (sigma1_sq + sigma2_sq + c2)
    ssim_map = ((2 * mu1_mu2 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1), torch.flatten(cs_map, 2).mean(-1)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 255.0) -> torch.Tensor:
    win = _window().to(x.dtype)
    weights = torch.tensor(WEIGHTS, dtype=x.dtype)
    mcs = []
    for i in range(5):
        ssim_per_channel, cs = _ssim(x, y, data_range, win)
        if i < 4:
            mcs.append(torch.relu(cs))
            padding = [s % 2 for s in x.shape[2:]]
            x, y = F.avg_pool2d(x, kernel_size=2, padding=padding), F.avg_pool2d(y, kernel_size=2, padding=padding)
    stacked = torch.stack(mcs + [torch.relu(ssim_per_channel)], dim=0)
    val = torch.prod(stacked ** weights.view(-1, 1, 1), dim=0)
    return val.mean(1).mean()


def average_metrics(tasks, x: Dict[str, torch.Tensor], x_hats: Dict[str, torch.Tensor], log_dir: str) -> Dict[str, torch.Tensor]:
    """mtc.py:359-384, metric by metric, task by task."""
    logs = {}
    with torch.no_grad():
        for name, fn in (("psnr", peak_signal_noise_ratio), ("ms-ssim", ms_ssim)):
            for task in tasks:
                pred, target = x_hats[task].detach(), x[task]
                if task == "semantic":
                    mult, data_range = 1, 17
                    pred = torch.argmax(pred, dim=1).unsqueeze(1).float()
                else:
                    mult, data_range = 255, 255
                logs[f"{log_dir}/{task}/{name}"] = fn(pred * mult, target * mult, data_range=data_range)
    return logs
