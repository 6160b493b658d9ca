"""C5 alone: the rANS sweep of bench.py (batch 1 ... 1024 at the C2 latent shapes).  usage: python tools/rans_bench.py"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mmnc_b200 as mm
dev = torch.device("cuda:0")
c2 = bench.CONFIGS["C2"]
m = mm.build_compressor(c2["model_type"], c2["tasks"], c2["latent_channels"], c2["conv_channels"])
m.update_bottleneck_values(); m.to(dev)
r = bench.rans_sweep(torch, mm, m, dev, 1, 0, None, cpu="--cpu" in sys.argv)
for row in r["sweep"]:
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items()})
print({k: v for k, v in r.items() if k not in ("sweep", "what")})
