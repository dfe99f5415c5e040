"""A few device-resident fp32 forwards of config C1 (784-128-64-10, 64 samples): the command profiled for gemm_fp32_ordered_small_kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
rng = np.random.default_rng(0)
npl, n_ins = [128, 64, 10], 784
w = rng.uniform(-1, 1, 784 * 128 + 128 * 64 + 64 * 10).astype(np.float32); b = rng.uniform(-1, 1, sum(npl)).astype(np.float32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_FP32); net.upload_mlp(w, b)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand((batch, n_ins), device="cuda") * 2 - 1; y = torch.empty((batch, 10), device="cuda")
for _ in range(6): net.forward_device(x, y, batch, s)
s.synchronize()
print("ok", float(y.abs().max()))
