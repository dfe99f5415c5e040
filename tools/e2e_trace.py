"""Host timeline of one net_cuda::launch_forward(std::vector) call on ViT-B/16-224, 1024 images (NETCUDA_HOST_TRACE)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
flat = nc.vit_random_params(cfg, seed=0)
B = 1024
x = np.random.default_rng(0).uniform(-1, 1, (B, 3 * 224 * 224)).astype(np.float32)
net = nc.HostNet.vit(cfg, flat, max_batch=512)
dt, y = net.time_launch_forward(x, reps=2)
print(f"{B / dt:.0f} images/s, {dt * 1e3:.2f} ms per call", flush=True)
import torch
px = torch.from_numpy(x).pin_memory()
py = torch.empty((B, 1000)).pin_memory()
n2 = nc.Net.vit(cfg, max_batch=512); n2.upload_vit(flat)
import time
for _ in range(2): n2.forward_into(px, py)
t = time.perf_counter(); n2.forward_into(px, py); print("pinned blocking ms", (time.perf_counter() - t) * 1e3)
o = np.empty((B, 1000), np.float32)
for _ in range(2): n2.forward(x)
t = time.perf_counter(); n2.forward(x); print("pageable numpy netcuda_forward ms", (time.perf_counter() - t) * 1e3)
