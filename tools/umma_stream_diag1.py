import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import netcuda as nc
from oracle import Oracle
o = Oracle()
rng = np.random.default_rng(1)
cases = (([32], 128, 64), ([32], 512, 64), ([32], 1024, 64), ([32], 2048, 64), ([32, 32], 128, 64), ([64], 640, 128), ([272, 48, 10], 1040, 64), ([4096] * 2, 4096, 128))
npl, n_ins, batch = cases[int(sys.argv[1])]
n_w = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
wq = rng.integers(-20, 20, n_w, dtype=np.int8); bq = rng.integers(-2000, 2000, sum(npl), dtype=np.int32)
xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
want = o.mlp_forward_i8(xq, wq, bq, npl, n_ins)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=160); net.upload_mlp_i8(wq, bq)
try:
    got = net.forward_i8(xq)
    print(npl, n_ins, batch, "mismatches", int((got != want).sum()), flush=True)
except Exception as e:
    print(npl, n_ins, batch, "ERROR", str(e)[:100], flush=True)
