"""Config C5 at 1..32 samples: the weight-streaming kernel with the tagged-word activation exchange (default up to 4 samples) against
the grid-barrier exchange (NETCUDA_MLP_STREAM_LL=0); us per forward and TB/s of weights, device-resident, bit-compared to each other."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
import vit_presets as vp
npl, n_ins = vp.MLP_C5["npl"], vp.MLP_C5["n_ins"]
wq, bq = vp.mlp_int8_params(npl, n_ins)
wbytes = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
batches = (1, 2, 4, 8, 9, 12, 16, 17, 32)
xs = {b: torch.randint(-128, 128, (b, n_ins), device="cuda", dtype=torch.int32).to(torch.int8) for b in batches}
outs = {}
variants = [(f"{'tagged words' if ll else 'grid barrier'}, L2 prefetch {pf}", {"NETCUDA_MLP_STREAM_LL": str(ll), "NETCUDA_MLP_STREAM_PF": str(pf)})
            for pf in (int(a) for a in (sys.argv[1:] or ["0", "6"])) for ll in (1, 0)]
for name, env in variants:
    os.environ.update(env)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=64)
    net.upload_mlp_i8(wq, bq)
    s = torch.cuda.Stream()
    row = []
    for batch in batches:
        xq = xs[batch]
        yq = torch.empty((batch, npl[-1]), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            for _ in range(20): net.forward_device_i8(xq, yq, batch, s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(200): net.forward_device_i8(xq, yq, batch, s)
            e1.record(s)
        s.synchronize()
        us = e0.elapsed_time(e1) * 5
        row.append(f"{batch}: {us:.1f} ({wbytes / us / 1e6:.2f})")
        ref = outs.setdefault(batch, yq.cpu())
        assert torch.equal(ref, yq.cpu()), (name, batch)
    print(f"{name:32s} us per forward (TB/s)  " + "  ".join(row), flush=True)
    net.close()
print("all exchanges bit-identical")
