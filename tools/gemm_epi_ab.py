"""A/B: the same ViT-B fc1-shaped GEMM (50432 x 3072 x 768, bf16 out) with different epilogues; CUDA-event time per launch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
import ctypes as C
m, k = 256 * 197, 768
for n in (3072, 2304):
    a = torch.randn((m, k), device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    o = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    s = torch.cuda.Stream(); torch.cuda.synchronize()
    for name, epi in (("none", nc.EPI_NONE), ("relu", nc.EPI_RELU), ("gelu", nc.EPI_GELU)):
        with torch.cuda.stream(s):
            for _ in range(3):
                nc.op_gemm(a, w, b, o, nc.PREC_BF16, nc.OUT_BF16, epilogue=epi, stream=s)
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                nc.op_gemm(a, w, b, o, nc.PREC_BF16, nc.OUT_BF16, epilogue=epi, stream=s)
                e1.record(s)
                s.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(f"N={n} epilogue={name}: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.0f} TFLOP/s", flush=True)
