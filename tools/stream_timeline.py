"""globaltimer timeline of the INT8 weight-streaming kernel (mlp_stream.cu), config C5, one CTA: per layer
[start, barrier passed, activations in registers, tiles done, partials visible, outputs released]."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(32 * 6, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_STREAM_DEBUG_PTR"] = hex(dbg.data_ptr())
os.environ["NETCUDA_STREAM_DEBUG_CTA"] = sys.argv[1] if len(sys.argv) > 1 else "0"
import netcuda as nc
np.set_printoptions(linewidth=200)
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=64); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for batch in (1, 4, 16):
    x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
    for _ in range(5): net.forward_device_i8(x, y, batch, s)
    s.synchronize()
    d = dbg.cpu().numpy().reshape(32, 6)[:8]
    extra = dbg.cpu().numpy().reshape(32, 6)[16:24, :2]
    print("batch", batch, "ns relative to layer 0 start; columns: start, barrier passed, act loaded, tiles done, partials visible, released")
    print(d - d[0, 0])
    if extra.any(): print("tagged-word finalize: [before store, after store] relative to 'partials visible'\n", (extra - d[:, 4:5]) * (extra != 0))
