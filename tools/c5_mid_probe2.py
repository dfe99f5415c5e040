"""Config C5 at 33..128 samples: the tcgen05 streaming kernel vs the split-K GEMM path vs the plain per-layer GEMMs (us per forward)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
import vit_presets as vp
npl, n_ins = vp.MLP_C5["npl"], vp.MLP_C5["n_ins"]
wq, bq = vp.mlp_int8_params(npl, n_ins)
for name, env, variant in (("tcgen05 stream", {}, 0), ("split-K GEMMs (graph)", {"NETCUDA_MLP_STREAM": "0"}, 0), ("plain GEMMs, single CTAs (graph)", {"NETCUDA_MLP_STREAM": "0"}, 2)):
    for k in ("NETCUDA_MLP_STREAM",): os.environ.pop(k, None)
    os.environ.update(env)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=256)
    net.upload_mlp_i8(wq, bq)
    if variant: net.set_gemm_variant(variant)
    s = torch.cuda.Stream()
    row = []
    for batch in (33, 64, 96, 128, 160, 256):
        xq = torch.randint(-128, 128, (batch, n_ins), device="cuda", dtype=torch.int32).to(torch.int8)
        yq = torch.empty((batch, npl[-1]), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            for _ in range(10): net.forward_device_i8(xq, yq, batch, s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(100): net.forward_device_i8(xq, yq, batch, s)
            e1.record(s)
        s.synchronize()
        row.append(f"{batch}: {e0.elapsed_time(e1) * 10:.1f}")
    print(f"{name:36s} us per forward  " + "  ".join(row), flush=True)
    net.close()
