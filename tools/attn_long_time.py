"""Time of the key-blocked tcgen05 attention kernel alone: ViT-L/16-384 shapes (32 images x 16 heads x 577 tokens)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
batch, tokens, heads = 32, 577, 16
qkv = torch.randn((batch * tokens, 3 * heads * 64), device="cuda").to(torch.bfloat16)
out = torch.empty((batch * tokens, heads * 64), dtype=torch.bfloat16, device="cuda")
for _ in range(5): nc.op_attention(qkv, out, batch, tokens, heads)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): nc.op_attention(qkv, out, batch, tokens, heads)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"us per launch {us:.1f} -> {4.0 * batch * heads * tokens * tokens * 64 / us / 1e6:.0f} TFLOP/s")
