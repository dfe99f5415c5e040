"""clock64 timeline of the tcgen05 INT8 streaming kernel (mlp_umma_stream.cu), config C5, CTA 0: per layer
[barrier passed, first operands landed, last MMA issued, accumulator complete, outputs published].  NETCUDA_DEBUG_TIMELINE build."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_STREAM_DEBUG_PTR"] = hex(dbg.data_ptr())
os.environ["NETCUDA_MLP_STREAM_SPLIT"] = "0"
import netcuda as nc
np.set_printoptions(linewidth=200)
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=128); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for batch in (8, 128):
    x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
    for _ in range(5): net.forward_device_i8(x, y, batch, s)
    s.synchronize()
    d = dbg.cpu().numpy().reshape(16, 8)[:8, :5]
    print("batch", batch, "cycles relative to layer 0's first operands; columns: barrier passed, first operands, last MMA issued, accumulator complete, published")
    print(d - d[0, 1])
