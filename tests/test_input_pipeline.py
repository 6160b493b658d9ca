"""(f4) GPU input pipeline against the reference's host transforms (restated with numpy / torch CPU ops):
ToTensor (/255), 16-bit depth scaling, semantic class remapping - bit for bit."""
import numpy as np
import pytest
import torch

import mmnc_b200 as mm
from mmnc_b200 import input_pipeline as ip


def _reference_transform(task, raw):
    """What get_transform + CLEVRDataset.__getitem__ produce for ONE sample (numpy in, torch CPU out)."""
    if task in ("rgb", "normal"):
        x = torch.from_numpy(raw.transpose(2, 0, 1)).contiguous().to(torch.float32).div(255)  # ToTensor on uint8 HWC
        return x[:3]
    if task == "depth_euclidean":
        return torch.from_numpy(raw.astype(np.int32))[None].float() / (2 ** 15 - 1.0)
    if task == "semantic":
        x = torch.from_numpy(raw.transpose(2, 0, 1)).contiguous()[1].clone()  # channel G: material, color
        for i, cls in enumerate(ip.SEM1_CLASSES):
            x[x == cls] = i
        return x.unsqueeze(0).float()
    raise NotImplementedError(task)


def _raw_batches(n_batches, B, tasks, seed=0):
    rng = np.random.default_rng(seed)
    for _ in range(n_batches):
        out = {}
        for t in tasks:
            if t == "depth_euclidean":
                out[t] = rng.integers(0, 2 ** 16, (B, 64, 48), dtype=np.uint16)
            elif t == "semantic":
                a = rng.integers(0, 256, (B, 64, 48, 3), dtype=np.uint8)
                a[..., 1] = rng.choice(np.array(ip.SEM1_CLASSES + (8, 9, 200), dtype=np.uint8), (B, 64, 48))
                out[t] = a
            else:
                out[t] = rng.integers(0, 256, (B, 64, 48, 4 if t == "rgb" else 3), dtype=np.uint8)
        yield out


def test_semantic_table_reproduces_the_in_place_loop():
    lut = ip.semantic_lut("cpu").numpy()
    x = torch.arange(256, dtype=torch.uint8)
    for i, cls in enumerate(ip.SEM1_CLASSES):
        x[x == cls] = i
    assert np.array_equal(lut, x.float().numpy())
    assert lut[255] == 16 and lut[17] == 15 and lut[10] == 8 and lut[8] == 8


@pytest.mark.gpu
def test_gpu_batch_loader_matches_reference_transforms():
    tasks = ("rgb", "depth_euclidean", "normal", "semantic")
    batches = list(_raw_batches(3, 5, tasks, seed=1))
    loader = ip.GpuBatchLoader(batches, tasks, device="cuda:0")
    n = 0
    for raw, got in zip(batches, loader):
        for t in tasks:
            want = torch.stack([_reference_transform(t, raw[t][b]) for b in range(raw[t].shape[0])])
            assert got[t].dtype == torch.float32 and got[t].shape == want.shape
            assert torch.equal(got[t].cpu(), want), t
        n += 1
    assert n == 3
    f32_bytes = sum(int(np.prod(got[t].shape)) * 4 for t in tasks)
    assert loader.h2d_bytes_per_batch * 2.5 < f32_bytes  # 12 raw bytes per pixel (4 + 2 + 3 + 3) against 32 as fp32
    with pytest.raises(TypeError):
        ip.convert_raw("rgb", torch.zeros(1, 4, 4, 3, device="cuda:0"))
