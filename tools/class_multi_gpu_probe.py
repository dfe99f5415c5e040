"""cuda::net_cuda with NETCUDA_DEVICES GPUs in one process: images/s of launch_forward-style calls (pageable host vectors in, vector out)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
flat = nc.vit_random_params(cfg, seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = np.random.default_rng(0).uniform(-1, 1, (B, 3 * 224 * 224)).astype(np.float32)
ref = None
for ndev in (1, 2, 4, 8):
    if ndev > nc.device_count(): break
    os.environ["NETCUDA_DEVICES"] = str(ndev)
    net = nc.HostNet.vit(cfg, flat, max_batch=512)
    dt, y = net.time_launch_forward(x, reps=3)
    if ref is None: ref = y
    print(f"{ndev} GPU(s) in one process: {B / dt:.0f} images/s through net_cuda::launch_forward (std::vector in, std::vector out), bit-equal to 1 GPU: {bool((y == ref).all())}", flush=True)
    net.close()
