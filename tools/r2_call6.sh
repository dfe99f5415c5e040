#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x -k "int8" --timeout 600 -p no:cacheprovider 2>&1 | tail -6
timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0, "vit-fpga_b200")
import numpy as np, torch
import netcuda as nc
import vit_presets as vp
npl, n_ins = vp.MLP_C5["npl"], vp.MLP_C5["n_ins"]
wq, bq = vp.mlp_int8_params(npl, n_ins)
import os
for split in ("32", "16", "8", "0"):
    os.environ["NETCUDA_MLP_STREAM_SPLIT"] = split
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=256)
    net.upload_mlp_i8(wq, bq)
    s = torch.cuda.Stream()
    row = []
    for batch in (1, 8, 16, 24, 32, 48, 64, 96, 128, 129):
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
    print(f"hand-over at {split}: us per forward  " + "  ".join(row), flush=True)
    net.close()
PY
