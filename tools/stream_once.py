"""A few INT8 8 x 4096 forwards of one sample (the weight-streaming kernel) -- the command profiled by tools/gpu_ncu_stream.sh."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=256); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
for _ in range(6): net.forward_device_i8(x, y, batch, s)
s.synchronize()
print("ok", int(y.abs().max()))
