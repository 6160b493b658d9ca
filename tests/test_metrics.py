"""(f2) Step metrics: MS-SSIM / PSNR restatement against an independent float64 numpy / scipy implementation (CPU), and
the fused argmax + squared-error kernel and the wrapper's metric logs on the GPU."""
import numpy as np
import pytest
import scipy.ndimage
import torch

import mmnc_b200 as mm
from mmnc_b200 import metrics as M


def _ms_ssim_np(x, y, data_range):
    """pytorch_msssim.ms_ssim written with scipy.ndimage in float64 ('valid' separable Gaussian, 5 scales)."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    coords = np.arange(11) - 5
    g = np.exp(-(coords ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()

    def filt(a):  # valid correlation along H then W, per (b, c) plane
        a = scipy.ndimage.correlate1d(a, g, axis=2, mode="constant")[:, :, 5:-5, :]
        return scipy.ndimage.correlate1d(a, g, axis=3, mode="constant")[:, :, :, 5:-5]

    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    weights = np.array(M.MS_SSIM_WEIGHTS)
    terms = []
    for i in range(5):
        mu1, mu2 = filt(x), filt(y)
        s1, s2, s12 = filt(x * x) - mu1 ** 2, filt(y * y) - mu2 ** 2, filt(x * y) - mu1 * mu2
        cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
        ssim_map = (2 * mu1 * mu2 + c1) / (mu1 ** 2 + mu2 ** 2 + c1) * cs_map
        cs, ss = cs_map.mean(axis=(2, 3)), ssim_map.mean(axis=(2, 3))
        if i < 4:
            terms.append(np.maximum(cs, 0))
            B, C, H, W = x.shape
            x, y = (a[:, :, :H // 2 * 2, :W // 2 * 2].reshape(B, C, H // 2, 2, W // 2, 2).mean(axis=(3, 5)) for a in (x, y))
    terms.append(np.maximum(ss, 0))
    return float(np.prod(np.stack(terms) ** weights[:, None, None], axis=0).mean())


def test_ms_ssim_and_psnr_against_float64_numpy():
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 256, 256)).astype(np.float32)
    for noise in (0.02, 0.2):
        y = np.clip(x + rng.standard_normal(x.shape).astype(np.float32) * noise, 0, 1)
        got = M.ms_ssim(torch.from_numpy(x) * 255, torch.from_numpy(y) * 255, data_range=255.0).item()
        want = _ms_ssim_np(x * 255, y * 255, 255.0)
        assert abs(got - want) <= 2e-5 * abs(want), (got, want)
        mse = torch.tensor(float(((x - y) ** 2).mean()))
        want_psnr = 10 * np.log10(255.0 ** 2 / (((x.astype(np.float64) - y) * 255) ** 2).mean())
        assert abs(M.psnr_from_mse(mse, 255.0, 255.0).item() - want_psnr) < 1e-3
    assert abs(M.ms_ssim(torch.from_numpy(x), torch.from_numpy(x), data_range=1.0).item() - 1.0) < 1e-6
    with pytest.raises(ValueError):
        M.ms_ssim(torch.zeros(1, 1, 128, 128), torch.zeros(1, 1, 128, 128))


def test_oracle_metric_restatement_against_float64_numpy():
    """oracle/metrics_ref.py (what the CPU reference arm runs and the GPU metrics are checked against) is pinned to the
    independent float64 implementation above; PSNR to its definition."""
    from oracle import metrics_ref
    rng = np.random.default_rng(1)
    x = rng.random((2, 3, 192, 224)).astype(np.float32)
    y = np.clip(x + rng.standard_normal(x.shape).astype(np.float32) * 0.1, 0, 1)
    got = metrics_ref.ms_ssim(torch.from_numpy(x) * 255, torch.from_numpy(y) * 255, data_range=255).item()
    want = _ms_ssim_np(x * 255, y * 255, 255.0)
    assert abs(got - want) <= 2e-5 * abs(want), (got, want)
    psnr = metrics_ref.peak_signal_noise_ratio(torch.from_numpy(x) * 255, torch.from_numpy(y) * 255, 255).item()
    assert abs(psnr - 10 * np.log10(255.0 ** 2 / (((x.astype(np.float64) - y) * 255) ** 2).mean())) < 1e-3


@pytest.mark.gpu
def test_semantic_argmax_sse_kernel_and_metric_logs():
    torch.manual_seed(3)
    dev = "cuda:0"
    logits = torch.randn(3, 17, 64, 48, device=dev)
    logits[0, 5, 0, 0] = logits[0, 9, 0, 0] = 50.0  # tie: the first maximum wins, like torch.argmax
    target = torch.randint(0, 17, (3, 1, 64, 48), device=dev).float()
    labels, sse = M.semantic_labels_and_sse(logits, target)
    want = torch.argmax(logits, dim=1).unsqueeze(1).float()
    assert torch.equal(labels, want) and labels[0, 0, 0, 0] == 5
    assert abs(sse.item() - ((want - target) ** 2).sum().item()) <= 1e-5 * sse.item()
    # the wrapper logs the reference's metric names on validation steps, and PSNR agrees with the definition
    tasks = ("rgb", "depth_euclidean", "semantic")
    model = mm.build_compressor(2, tasks, 12, 8, lmbda=1e-2).to(dev)
    batch = mm.synthetic_batch(tasks, 2, size=256, seed=4, device=dev)
    model.validation_step(batch)
    logs = model.last_logs
    for t in tasks:
        assert f"val/{t}/psnr" in logs and f"val/{t}/ms-ssim" in logs
    with torch.no_grad():
        x_hats, _ = model.eval()(batch)
    mse = ((x_hats["rgb"] - batch["rgb"]) ** 2).mean()
    assert abs(logs["val/rgb/psnr"].item() - (-10 * torch.log10(mse)).item()) < 1e-3
    assert 0.0 <= logs["val/rgb/ms-ssim"].item() <= 1.0
    # every metric of the validation step against the oracle's restatement of torchmetrics / pytorch_msssim on the CPU
    from oracle import metrics_ref
    want = metrics_ref.average_metrics(tasks, {k: v.cpu() for k, v in batch.items()}, {k: v.cpu() for k, v in x_hats.items()}, "val")
    for k, v in want.items():
        assert abs(logs[k].item() - v.item()) <= 1e-4 * max(1.0, abs(v.item())), (k, logs[k].item(), v.item())
    model.train()
    model.configure_optimizers(total_steps=4)
    model.training_step(batch)   # like the reference (mtc.py:468): metrics on every training step too
    assert "train/rgb/psnr" in model.last_logs and "train/semantic/ms-ssim" in model.last_logs
    model.train_metrics_every = 0
    model.training_step(batch)
    assert not any(k.endswith("psnr") for k in model.last_logs)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 256, 256), (3, 1, 176, 208), (1, 2, 161, 170), (5, 3, 192, 256)])
def test_ms_ssim_kernel_against_float64_numpy_and_torch_ops(shape):
    """mmnc_ssim_scale (one fused launch per scale, pooled next-scale input written by the same block) against the
    independent float64 numpy implementation (2e-5 relative, the bar of the CPU test above) and against the stock-torch
    restatement on the same device; 176 x 208: tiles that end inside the image and an 11 x 13 last scale; 161 x 170:
    odd sides, where pytorch_msssim's zero-padded pooling is done with torch ops between the launches."""
    rng = np.random.default_rng(5)
    x = rng.random(shape).astype(np.float32)
    dev = "cuda:0"
    for noise in (0.02, 0.2):
        y = np.clip(x + rng.standard_normal(shape).astype(np.float32) * noise, 0, 1)
        xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
        before = mm.launch_count()
        got = M.ms_ssim(xd, yd, data_range=255.0, scale=255.0)
        torch.cuda.synchronize()
        assert mm.launch_count() - before == 10  # 5 scales x (tile kernel + fixed-order finish)
        ref_t = M.ms_ssim_torch(xd * 255, yd * 255, data_range=255.0).item()
        assert abs(got.item() - ref_t) <= 2e-5 * abs(ref_t), (got.item(), ref_t)
        if shape[2] % 16 == 0 and shape[3] % 16 == 0:  # the numpy twin crops instead of padding odd sides
            want = _ms_ssim_np(x * 255, y * 255, 255.0)
            assert abs(got.item() - want) <= 2e-5 * abs(want), (got.item(), want)
        per_image = M.ms_ssim(xd, yd, data_range=255.0, scale=255.0, size_average=False)
        assert per_image.shape == (shape[0],)
        assert torch.allclose(per_image, M.ms_ssim_torch(xd * 255, yd * 255, 255.0, size_average=False), rtol=2e-5)
        assert torch.equal(M.ms_ssim(xd, yd, data_range=255.0, scale=255.0), got), "fixed-order reductions: bit-reproducible"
    same = M.ms_ssim(torch.from_numpy(x).to(dev), torch.from_numpy(x).to(dev), data_range=1.0)
    assert abs(same.item() - 1.0) < 1e-6
