"""clock64 timeline of the CTA-pair GEMM (cluster 0): MMA issuer and epilogue warps, ViT-B fc1 / qkv shapes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(40 * 24 * 4, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_GEMM_DEBUG_PTR"] = hex(dbg.data_ptr())
import netcuda as nc
m, k = 256 * 197, 768
np.set_printoptions(linewidth=250)
for name, n, epi in (("fc1 gelu", 3072, nc.EPI_GELU), ("qkv", 2304, nc.EPI_NONE)):
    a = torch.randn((m, k), device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    o = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        dbg.zero_()
        nc.op_gemm(a, w, b, o, nc.PREC_BF16, nc.OUT_BF16, epilogue=epi)
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().reshape(40, 24, 4)
    t0 = d[0, 11, 0]
    print("====", name)
    print("MMA issuer: wait tempty start, tempty ready, last commit issued  (per tile, rel cycles)")
    print((d[:24, 11, :3] - t0).T)
    for wv in (0, 4, 12 + 0, 12 + 7):
        print(f"epilogue warp {wv % 12} of CTA {wv // 12}: wait tfull start, tfull ready, reads done")
        print((d[:24, wv, :3] - t0).T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nc.op_gemm(a, w, b, o, nc.PREC_BF16, nc.OUT_BF16, epilogue=epi)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"us per launch {us:.1f}  -> {2.0 * m * n * k / us / 1e6:.0f} TFLOP/s")
