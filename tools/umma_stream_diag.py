"""Diagnostics of the tcgen05 INT8 streaming kernel: small nets against numpy, mismatch patterns."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import netcuda as nc
from oracle import Oracle
o = Oracle()
rng = np.random.default_rng(1)
for npl, n_ins, batch in (([32], 128, 33), ([32], 256, 40), ([64], 128, 128), ([32], 48, 33), ([10], 128, 33), ([32, 32], 128, 33), ([272, 48, 10], 1040, 33), ([4096] * 2, 4096, 128)):
    n_w = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = rng.integers(-20, 20, n_w, dtype=np.int8)
    bq = rng.integers(-2000, 2000, sum(npl), dtype=np.int32)
    xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
    want = o.mlp_forward_i8(xq, wq, bq, npl, n_ins)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=160)
    net.upload_mlp_i8(wq, bq)
    net.profile_enable(True)
    got = net.forward_i8(xq)
    labels = list(net.profile_read())
    bad = got != want
    print(npl, n_ins, batch, labels, "mismatches", int(bad.sum()), "of", bad.size,
          "| bad rows", np.unique(np.nonzero(bad)[0])[:12], "| bad cols", np.unique(np.nonzero(bad)[1])[:12], flush=True)
    if bad.any():
        r, c = np.argwhere(bad)[0]
        print("   first:", r, c, "got", got[r, c], "want", want[r, c])
    net.close()
