import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
flat = nc.vit_random_params(cfg, seed=0)
x = torch.rand((1024, 3 * 224 * 224), device="cuda") * 2 - 1
outs = {}
for mb in (256, 1024):
    net = nc.Net.vit(cfg, max_batch=mb)
    net.upload_vit(flat)
    y = torch.empty((1024, 1000), device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        net.forward_device(x, y, 1024, s)
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(s)
        for _ in range(10):
            net.forward_device(x, y, 1024, s)
        e1.record(s)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"mb={mb} rep={rep}: events {e0.elapsed_time(e1)/10:.2f} ms/step, wall {1e3*(t2-t0)/10:.2f} ms/step, launch-side {1e3*(t1-t0)/10:.2f} ms/step")
    outs[mb] = y.clone()
    net.close()
print("equal:", torch.equal(outs[256], outs[1024]), "max diff", (outs[256] - outs[1024]).abs().max().item(), "absmax", outs[256].abs().max().item())
