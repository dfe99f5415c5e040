"""A/B of GEMM variants inside the ViT-B step: bit-equality of the logits against variant 0 and per-kernel times."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
variants = [int(v) for v in sys.argv[1:]] or [0, 4]
cfg = nc.VIT_PRESETS["vit_base_16_224"]
B = 1024
net = nc.Net.vit(cfg, max_batch=512)
net.upload_vit(nc.vit_random_params(cfg, seed=0))
x = torch.rand((B, net.n_in), device="cuda") * 2 - 1
y = torch.empty((B, net.n_out), device="cuda")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ref = None
for rep in range(2):
    for v in variants:
        net.set_gemm_variant(v)
        for _ in range(3): net.forward_device(x, y, B, s)
        s.synchronize()
        if ref is None: ref = y.clone()
        same = bool((y == ref).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(15): net.forward_device(x, y, B, s)
        e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / 15
        net.profile_enable(True)
        for _ in range(2): net.forward_device(x, y, B, s)
        s.synchronize(); prof = net.profile_read(); net.profile_enable(False)
        print(f"variant {v}: {ms:.2f} ms/step {B / ms * 1e3:.0f} img/s  bit-equal to variant {variants[0]}: {same}  ",
              {k: round(p["ms"] / 2, 2) for k, p in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:7]}, flush=True)
