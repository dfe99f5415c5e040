"""Single-sample latency (the reference's launch_forward contract: one sample per call): device-resident and host-buffer calls."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name in ("vit_tiny_16_224", "vit_base_16_224"):
    cfg = nc.VIT_PRESETS[name]
    net = nc.Net.vit(cfg, max_batch=8)
    net.upload_vit(nc.vit_random_params(cfg, seed=0))
    for batch in (1, 8):
        x = torch.rand((batch, net.n_in), device="cuda") * 2 - 1; y = torch.empty((batch, net.n_out), device="cuda")
        for _ in range(5): net.forward_device(x, y, batch, s)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(50): net.forward_device(x, y, batch, s)
        e1.record(s); s.synchronize()
        hx = x.cpu().pin_memory(); hy = torch.empty((batch, net.n_out)).pin_memory()
        for _ in range(3): net.forward_into(hx, hy)
        t0 = time.perf_counter()
        for _ in range(50): net.forward_into(hx, hy)
        host_us = (time.perf_counter() - t0) / 50 * 1e6
        print(f"{name} batch {batch}: device-resident {e0.elapsed_time(e1) / 50 * 1e3:.0f} us per call, host-buffer call {host_us:.0f} us; launches per call {net.launches // (5 + 50 + 3 + 50) if batch == 1 else ''}")
    net.close()
