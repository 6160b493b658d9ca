"""Where the end-to-end training step goes (torch.profiler, top device kernels).  usage: python tools/e2e_profile.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
TASKS = ("rgb", "depth_euclidean", "normal")
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(21)
model = mm.build_compressor(3, TASKS, 128, 100, lmbda=1e-2)
model.update_bottleneck_values(); model.to(dev); model.train()
if "channels_last" in sys.argv: model.use_channels_last()
model.configure_optimizers(total_steps=100)
host = mm.synthetic_batch(TASKS, 64, seed=21, pin_memory=True)
def step():
    b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    return float(model.training_step(b).item())
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(e.device_time_total for e in ka if e.device_type == torch.autograd.DeviceType.CUDA) if hasattr(torch.autograd, "DeviceType") else 0
rows = sorted(ka, key=lambda e: -e.self_device_time_total)[:28]
tot = sum(e.self_device_time_total for e in ka)
print("total device time per step (ms):", tot / 3 / 1e3)
for e in rows:
    print(f"{e.self_device_time_total/3/1e3:8.3f} ms  x{e.count//3:4d}  {e.key[:110]}")
# which operators launch the memsets (device time per step), and how many bytes' worth at ~3 TB/s
by_op = {}
for e in prof.events():
    for k in getattr(e, "kernels", []) or []:
        if "emset" in k.name:
            d = by_op.setdefault(e.name, [0.0, 0])
            d[0] += k.duration
            d[1] += 1
print("memsets by launching op:")
for name, (dur, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"{dur/3/1e3:8.3f} ms  x{n//3:4d}  {name[:100]}")
