"""Which layers of the -m 3 model see / produce channels-last tensors after use_channels_last()."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmnc_b200 as mm
TASKS = ("rgb", "depth_euclidean", "normal")
dev = torch.device("cuda:0")
model = mm.build_compressor(3, TASKS, 128, 100, lmbda=1e-2).to(dev).train().use_channels_last()
def fmt(t):
    if not torch.is_tensor(t) or t.dim() != 4: return "-"
    a, b = t.is_contiguous(), t.is_contiguous(memory_format=torch.channels_last)
    return "both" if a and b else ("NCHW" if a else ("NHWC" if b else "strided"))
rows = []
for name, m in model.named_modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, mm.GDN)):
        m.register_forward_hook(lambda mod, i, o, name=name: rows.append((name, type(mod).__name__, tuple(i[0].shape), fmt(i[0]), fmt(o))))
batch = mm.synthetic_batch(TASKS, 2, device=dev)
model(batch)
for r in rows:
    if "heads.1" in r[0] or "heads.2" in r[0]: continue
    print(r)
