"""Stress of the tcgen05 INT8 streaming kernel (mlp_umma_stream.cu): many launches of ragged nets at varying batch sizes, each
checked against the oracle; prints which (batch, rows, columns) ever differ.  Diagnostics, not a test."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import netcuda as nc
from oracle import Oracle

o = Oracle()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for npl, n_ins in (([272, 48, 10], 1040), ([64, 32], 4080), ([4096, 304, 4096], 4096)):
    rng = np.random.default_rng(78)
    n_w = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = np.clip(np.rint(rng.standard_normal(n_w) * 128.0 / np.sqrt(n_ins) * 1.4), -128, 127).astype(np.int8)
    bq = rng.integers(-2000, 2000, sum(npl), dtype=np.int32)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, activation=0, max_batch=160)
    net.upload_mlp_i8(wq, bq)
    cases = {}
    for b in (64, 100, 127, 128, 65, 71, 96, 120, 127, 121):
        xq = rng.integers(-128, 128, (b, n_ins), dtype=np.int8)
        cases.setdefault(b, (xq, o.mlp_forward_i8(xq, wq, bq, npl, n_ins, 0)))
    order = list(cases)
    fails = {}
    for it in range(reps):
        for b in (order if it % 2 == 0 else order[::-1]):
            xq, want = cases[b]
            got = net.forward_i8(xq)
            bad = got != want
            if bad.any():
                rows, cols = np.unique(np.nonzero(bad)[0]), np.unique(np.nonzero(bad)[1])
                f = fails.setdefault(b, [0, None])
                f[0] += 1
                if f[1] is None:
                    f[1] = (it, rows[:6].tolist(), int(rows.size), cols.tolist()[:12], int(np.abs(got.astype(np.int64) - want)[bad].max()))
    print(npl, n_ins, "launches per batch", reps, "| failures:", {b: (f[0], f[1]) for b, f in fails.items()} or "none", flush=True)
    net.close()
