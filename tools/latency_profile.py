"""Per-kernel times of a single-sample ViT-B pass (profile mode: plain launches, one event pair per kernel)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name in ("vit_base_16_224",):
    cfg = nc.VIT_PRESETS[name]
    net = nc.Net.vit(cfg, max_batch=8)
    net.upload_vit(nc.vit_random_params(cfg, seed=0))
    for batch in (1, 8):
        x = torch.rand((batch, net.n_in), device="cuda") * 2 - 1; y = torch.empty((batch, net.n_out), device="cuda")
        for _ in range(5): net.forward_device(x, y, batch, s)
        s.synchronize()
        net.profile_enable(True)
        for _ in range(10): net.forward_device(x, y, batch, s)
        s.synchronize(); prof = net.profile_read(); net.profile_enable(False)
        print(name, "batch", batch, {k: (v["launches"] // 10, round(v["ms"] / v["launches"] * 1e3, 1)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}, "sum us", round(sum(v["ms"] for v in prof.values()) / 10 * 1e3))
