"""Per-phase clock64 timeline of the tcgen05 attention kernel (CTA 0), ViT-B shapes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(32 * 12 * 8, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_ATTENTION_DEBUG_PTR"] = hex(dbg.data_ptr())
import netcuda as nc
batch, tokens, heads = 256, 197, 12
qkv = torch.randn((batch * tokens, 3 * heads * 64), device="cuda").to(torch.bfloat16)
out = torch.empty((batch * tokens, heads * 64), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    nc.op_attention(qkv, out, batch, tokens, heads)
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(32, 12, 8)
t0 = d[0, 0, 0]
np.set_printoptions(linewidth=220)
print("producer issue (rel):", (d[:21, 0, 0] - t0))
print("mma S0,S1,PV0,PV1 per item:")
print(d[:21, 1, :4] - t0)
for w in (4, 8):
    print(f"warp {w}: wait_start, S ready, softmax start, softmax done(arrive), O ready, O in registers, stores done")
    print(d[:21, w, [0, 1, 2, 3, 4, 6, 5]] - t0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    nc.op_attention(qkv, out, batch, tokens, heads)
e1.record(); torch.cuda.synchronize()
print("us per launch", e0.elapsed_time(e1) / 20 * 1e3)
