"""BASELINE config C4: MultiTaskSharedLatentCompressor (-m 4), 4 tasks, eval bpp / likelihood pass at batch 256."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
tasks = ("rgb", "depth_euclidean", "normal", "semantic")
torch.manual_seed(21)
model = mm.build_compressor(4, tasks, 192, 128, lmbda=1e-2)
model.update_bottleneck_values(); model.to(dev).eval()
B = 256
batch = mm.synthetic_batch(tasks, B, device=dev, seed=21)
def step():
    with torch.no_grad():
        x_hats, lik = model(batch)
        loss, logs = model.rate_distortion_loss(batch, x_hats, lik, "val")
    return loss, logs
for _ in range(2): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): loss, logs = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print(f"C4 eval pass, batch {B}: {dt*1e3:.1f} ms ({B/dt:.0f} images/s), loss {loss.item():.4f}")
print({k: round(float(v), 5) for k, v in logs.items() if "bpp" in k or "compression" in k})
print("peak memory GB:", torch.cuda.max_memory_allocated() / 1e9)
