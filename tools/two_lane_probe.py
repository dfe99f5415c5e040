"""Probe: do two independent passes of ViT-B on two streams (own work buffers each) finish sooner than back to back on one?
HBM-bound kernels of one lane (LayerNorm, proj) can share the SMs with the other lane's tensor-bound GEMM CTAs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
params = nc.vit_random_params(cfg, seed=0)
for mb in (512, 256):
    nets = [nc.Net.vit(cfg, max_batch=mb) for _ in range(2)]
    for n in nets: n.upload_vit(params)
    x = torch.rand((1024, nets[0].n_in), device="cuda") * 2 - 1
    y = torch.empty((1024, 1000), device="cuda")
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    def seq():
        nets[0].forward_device(x, y, 1024, s0)
    def two():
        ev = torch.cuda.Event(); ev.record(s0); s1.wait_event(ev)
        nets[0].forward_device(x[:512], y[:512], 512, s0)
        nets[1].forward_device(x[512:], y[512:], 512, s1)
        ev2 = torch.cuda.Event(); ev2.record(s1); s0.wait_event(ev2)
    for name, fn in (("one lane", seq), ("two lanes", two), ("one lane", seq), ("two lanes", two)):
        for _ in range(3): fn()
        s0.synchronize(); s1.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s0)
        for _ in range(10): fn()
        e1.record(s0); s0.synchronize(); s1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"pass size {mb}: {name}: {ms:.2f} ms per 1024 images -> {1024 / ms * 1e3:.0f} images/s", flush=True)
    y2 = y.clone(); seq(); s0.synchronize()
    print("max abs diff two-lane vs one-lane logits:", float((y - y2).abs().max()))
    for n in nets: n.close()
