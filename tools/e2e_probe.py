"""Where the end-to-end step goes: H2D bandwidth alone / next to the kernels, device-resident step, pipelined host-buffer step."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
B, PASS = 1024, int(os.environ.get("PASS", "512"))
net = nc.Net.vit(cfg, max_batch=PASS)
net.upload_vit(nc.vit_random_params(cfg, seed=0))
n_in, n_out = net.n_in, net.n_out
hx = [torch.empty((B, n_in), dtype=torch.float32, pin_memory=True) for _ in range(2)]
hy = [torch.empty((B, n_out), dtype=torch.float32, pin_memory=True) for _ in range(2)]
hx[0].uniform_(-1, 1); hx[1].copy_(hx[0])
dx = torch.empty((B, n_in), device="cuda"); dy = torch.empty((B, n_out), device="cuda")
s_copy, s_comp = torch.cuda.Stream(), torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
# (a) H2D alone
for _ in range(2): dx.copy_(hx[0], non_blocking=True)
torch.cuda.synchronize()
a, b = ev(), ev(); a.record()
for _ in range(5): dx.copy_(hx[0], non_blocking=True)
b.record(); torch.cuda.synchronize()
print("H2D alone: %.1f GB/s" % (5 * B * n_in * 4 / a.elapsed_time(b) / 1e6))
# (b) device-resident step alone
for _ in range(3): net.forward_device(dx, dy, B, s_comp)
torch.cuda.synchronize()
a, b = ev(), ev(); a.record(s_comp)
for _ in range(10): net.forward_device(dx, dy, B, s_comp)
b.record(s_comp); torch.cuda.synchronize()
t_dev = a.elapsed_time(b) / 10
print("device step alone: %.2f ms" % t_dev)
# (c) both at once
dx2 = torch.empty_like(dx)
a, b, c, d = ev(), ev(), ev(), ev()
a.record(s_comp); c.record(s_copy)
for _ in range(10): net.forward_device(dx, dy, B, s_comp)
with torch.cuda.stream(s_copy):
    for _ in range(10): dx2.copy_(hx[0], non_blocking=True)
b.record(s_comp); d.record(s_copy); torch.cuda.synchronize()
print("concurrent: device step %.2f ms, H2D %.1f GB/s" % (a.elapsed_time(b) / 10, 10 * B * n_in * 4 / c.elapsed_time(d) / 1e6))
# (d) pipelined host-buffer calls
for depth in (2, 3):
    net.wait(net.submit(hx[0], hy[0])); torch.cuda.synchronize()
    t0 = time.perf_counter(); q = []
    for i in range(12):
        q.append(net.submit(hx[i & 1], hy[i & 1]))
        if len(q) >= depth: net.wait(q.pop(0))
    while q: net.wait(q.pop(0))
    dt = (time.perf_counter() - t0) / 12
    print("pipelined e2e, %d calls in flight: %.2f ms per step -> %.0f images/s" % (depth, dt * 1e3, B / dt))
t0 = time.perf_counter()
for i in range(6): net.forward_into(hx[0], hy[0])
dt = (time.perf_counter() - t0) / 6
print("synchronous e2e: %.2f ms per step -> %.0f images/s" % (dt * 1e3, B / dt))
