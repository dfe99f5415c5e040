#!/usr/bin/env python
"""Short-K GEMMs of a ViT-Tiny pass (256 images: M = 50432), one launch configuration against another:
variant 0 = product path, 3 = 16 epilogue warps, 4 = two slabs per epilogue warp, 5 = neither."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-fpga_b200"))
import torch
import netcuda as nc

M = 256 * 197
SHAPES = [("qkv", 576, 192, nc.OUT_BF16, nc.EPI_NONE), ("proj", 192, 192, nc.OUT_F32, nc.EPI_RESIDUAL), ("fc1", 768, 192, nc.OUT_BF16, nc.EPI_GELU),
          ("fc2", 192, 768, nc.OUT_F32, nc.EPI_RESIDUAL)]
MB = 512 * 197  # ... and the GEMMs of a ViT-B pass of 512 images
SHAPES = [(M,) + t for t in SHAPES] + [(MB, "B qkv", 2304, 768, nc.OUT_BF16, nc.EPI_NONE), (MB, "B proj", 768, 768, nc.OUT_F32, nc.EPI_RESIDUAL),
                                        (MB, "B fc1", 3072, 768, nc.OUT_BF16, nc.EPI_GELU), (MB, "B fc2", 768, 3072, nc.OUT_F32, nc.EPI_RESIDUAL)]
g = torch.Generator(device="cuda").manual_seed(1)
for M, name, n, k, out_type, epi in SHAPES:
    a = torch.randn((M, k), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, generator=g, device="cuda")
    outs = {}
    line = f"{name:6s} N={n:4d} K={k:4d}:"
    for variant in (0, 3, 4, 5):
        out = torch.zeros((M, n), dtype=torch.float32 if out_type == nc.OUT_F32 else torch.bfloat16, device="cuda")
        for _ in range(5):
            nc.op_gemm(a, w, b, out, nc.PREC_BF16, out_type, epilogue=epi, variant=variant)
        torch.cuda.synchronize()
        reps = 50
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            nc.op_gemm(a, w, b, out, nc.PREC_BF16, out_type, epilogue=epi, variant=variant)
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) / reps * 1e3
        out.zero_()
        nc.op_gemm(a, w, b, out, nc.PREC_BF16, out_type, epilogue=epi, variant=variant)
        torch.cuda.synchronize()
        outs[variant] = out.clone()
        line += f"  v{variant} {us:6.1f} us"
    same = all(torch.equal(outs[0], outs[v]) for v in (3, 4, 5))
    print(line, " identical bits" if same else " BITS DIFFER")
