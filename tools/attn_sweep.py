"""A/B of the short-sequence attention kernel's build variants (netcuda_op_attention_ex codes, csrc/attention.cu) on ViT-B / ViT-Tiny shapes:
CUDA-event time per launch (median of 15), all variants in one process on one box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
s = torch.cuda.Stream()
for batch, tokens, heads in ((512, 197, 12), (256, 197, 3)):
    qkv = torch.randn((batch * tokens, 3 * heads * 64), device="cuda").to(torch.bfloat16)
    out = torch.empty((batch * tokens, heads * 64), dtype=torch.bfloat16, device="cuda")
    ref = None
    torch.cuda.synchronize()
    for kernel in (0, 1, 2, 4, 14, 24, 34, 104, 4):
        with torch.cuda.stream(s):
            for _ in range(3):
                nc.op_attention_ex(qkv, out, batch, tokens, heads, kernel=kernel, stream=s)
            ts = []
            for _ in range(15):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                nc.op_attention_ex(qkv, out, batch, tokens, heads, kernel=kernel, stream=s)
                e1.record(s)
                s.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        o = out.float()
        if ref is None:
            ref = o.clone()
        err = ((o - ref).abs().max() / ref.abs().max()).item()
        flops = 4.0 * batch * heads * tokens * tokens * 64
        print(f"{batch}x{heads}x{tokens} kernel {kernel:2d}: {ts[len(ts) // 2]:7.1f} us (min {ts[0]:.1f})  {flops / ts[len(ts) // 2] / 1e6:6.1f} TFLOP/s  "
              f"max diff vs kernel 0: {err:.2e}", flush=True)
