import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo/vit-fpga_b200")
import netcuda as nc
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for variant in (0, 5):
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=2048); net.upload_mlp_i8(wq, bq)
    net.set_gemm_variant(variant)
    for batch in (16, 24, 32, 48, 160, 256, 512, 1024):
        x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
        for _ in range(5): net.forward_device_i8(x, y, batch, s)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): net.forward_device_i8(x, y, batch, s)
        e1.record(s); s.synchronize()
        print("variant", variant, "batch", batch, round(e0.elapsed_time(e1) / 20 * 1e3, 1), "us")
    net.close()
