"""Two compress / decompress passes at the C2 latent shapes, 1024 images, for ncu (profiles/)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mmnc_b200 as mm
dev = torch.device("cuda:0")
c2 = bench.CONFIGS["C2"]
m = mm.build_compressor(c2["model_type"], c2["tasks"], c2["latent_channels"], c2["conv_channels"])
m.update_bottleneck_values(); m.to(dev)
c = m.model["compressor"]; eb, gc = c.entropy_bottleneck, c.gaussian_conditional
B = 1024
g = torch.Generator(device=dev).manual_seed(5)
scales = torch.exp(torch.empty(B, c.M, 1, 1, device=dev).uniform_(-3.0, 4.16, generator=g))
y = torch.randn(B, c.M, 1, 1, device=dev, generator=g) * scales
z = torch.randn(B, c.N, 1, 1, device=dev, generator=g) * 4
for _ in range(2):
    idx = gc.build_indexes(scales)
    ys, zs = gc.compress(y, idx), eb.compress(z)
    yh, zh = gc.decompress(ys, idx), eb.decompress(zs, (1, 1))
torch.cuda.synchronize()
assert torch.equal(yh, torch.round(y))
print("ok", sum(map(len, ys)) + sum(map(len, zs)), "bytes")
