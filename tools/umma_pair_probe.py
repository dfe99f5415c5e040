"""A/B of the two tcgen05 INT8 streaming kernels on config C5 (8 x 4096, device-resident): split-K CTA clusters
(mlp_i8_umma_cluster_kernel; NETCUDA_MLP_UMMA_PAIR = 1: the default -- four CTAs per tile up to 88 samples, pairs above; 4: four CTAs per tile; 2: pairs, four issuers; 3: pairs, two issuers)
against single CTAs (NETCUDA_MLP_UMMA_PAIR=0).  Every timing is preceded by a bit-for-bit
comparison of the two kernels' outputs with each other (the parity tests compare both with the oracle).

    python tools/umma_pair_probe.py            # both kernels, batches 17..128
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BATCHES = (17, 32, 48, 64, 96, 128)

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "vit-fpga_b200"))
    import netcuda as nc
    rng = np.random.default_rng(0)
    npl, n_ins = [4096] * 8, 4096
    wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8)
    bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=256)
    net.upload_mlp_i8(wq, bq)
    g = torch.Generator(device="cuda").manual_seed(5)
    out = {}
    for batch in BATCHES:
        x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda", generator=g)
        y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
        net.profile_enable(True)
        net.forward_device_i8(x, y, batch, s)
        s.synchronize()
        labels = sorted(net.profile_read())
        net.profile_enable(False)
        for _ in range(10):
            net.forward_device_i8(x, y, batch, s)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(50):
            net.forward_device_i8(x, y, batch, s)
        e1.record(s)
        s.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        out[batch] = y.cpu().numpy()
        print(f"  batch {batch:4d}: {us:6.1f} us per forward  ({134217728 / us / 1e3:7.1f} GB/s of weights)  kernels {labels}", flush=True)
    np.savez(sys.argv[2], **{str(k): v for k, v in out.items()})
    net.close()
    sys.exit(0)

import numpy as np

files = {}
for pair in (1, 4, 2, 3, 0, 1):
    env = dict(os.environ, NETCUDA_MLP_UMMA_PAIR=str(pair))
    f = f"/tmp/umma_pair_{pair}.npz"
    print(f"NETCUDA_MLP_UMMA_PAIR={pair}", flush=True)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", f], env=env, timeout=600)
    if r.returncode != 0:
        print("  worker failed, rc", r.returncode)
        sys.exit(1)
    files[pair] = f
a, b = np.load(files[1]), np.load(files[0])
print("every cluster mode == single, bit for bit:", all(np.array_equal(np.load(files[m])[k], b[k]) for m in (1, 2, 3, 4) for k in b.files))
