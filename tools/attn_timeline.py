"""Per-phase clock64 timeline of the short-sequence tcgen05 attention kernel (CTA 0), ViT-B shapes.
Needs a NETCUDA_DEBUG_TIMELINE build:  NETCUDA_BUILD_TAG=dbg NETCUDA_DEBUG_TIMELINE=1 python vit-fpga_b200/build.py
                                        NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib_dbg python tools/attn_timeline.py [kernel ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(32 * 20 * 8, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_ATTENTION_DEBUG_PTR"] = hex(dbg.data_ptr())
import netcuda as nc
batch, tokens, heads = 256, 197, 12
qkv = torch.randn((batch * tokens, 3 * heads * 64), device="cuda").to(torch.bfloat16)
out = torch.empty((batch * tokens, heads * 64), dtype=torch.bfloat16, device="cuda")
np.set_printoptions(linewidth=220)
for kernel in [int(a) for a in sys.argv[1:]] or [0, 2]:
    dbg.zero_()
    for _ in range(3):
        nc.op_attention_ex(qkv, out, batch, tokens, heads, kernel=kernel)
    torch.cuda.synchronize()
    nw = 20 if kernel % 100 == 4 else 12
    d = dbg.cpu().numpy()[: 32 * nw * 8].reshape(32, nw, 8)
    lo, hi = 4, 19
    print(f"==== kernel {kernel}: mean cycles per item over items {lo}..{hi} of CTA 0")
    for w in ((4, 8, 7, 12, 16) if kernel % 100 == 4 else (4, 5, 7, 8, 10)):
        s = d[lo:hi + 1, w, :].astype(np.float64)
        period = np.diff(d[lo:hi + 2, w, 0]).mean()
        names = ["wait S", "pass1 (max)", "wait turn", "pass2 (exp2)", "wait O (P.V)", "read O", "store O"]
        order = [0, 1, 2, 7, 3, 4, 6, 5]
        if kernel % 10 != 2 and kernel % 100 != 4:  # (only MODE 2 stamps the turn; kernel 4: slot 7 = after the max exchange)
            s[:, 7] = s[:, 2]
        ph = [(s[:, order[i + 1]] - s[:, order[i]]).mean() for i in range(7)]
        print(f"warp {w:2d} (tile {(w - 4) // 4}, quarter {w % 4}): period {period:7.0f} | " + " | ".join(f"{n} {v:6.0f}" for n, v in zip(names, ph)))
    if kernel % 10 >= 1:
        for w in (1, 3):
            s = d[lo:hi + 1, w, :].astype(np.float64)
            print(f"issuer warp {w}: S issue -> P.V issue {np.mean(s[:, 2] - s[:, 0]):7.0f}; P.V issue -> next S issue {np.mean(d[lo + 1:hi + 2, w, 0] - d[lo:hi + 1, w, 2]):7.0f}")
    # lag between a softmax warp's arrive and the issuer's P.V issue, and between O read (sfree) and next S issue
    if kernel % 10 >= 1:
        t0w = [4, 5, 6, 7, 8, 9, 10, 11] if kernel % 100 == 4 else [4, 5, 6, 7]
        lagpv = (d[lo:hi + 1, 1, 2] - d[lo:hi + 1, t0w, 3].max(axis=1)).mean()
        lags = (d[lo + 1:hi + 2, 1, 0] - d[lo:hi + 1, t0w, 6].max(axis=1)).mean()
        print(f"issuer warp 1: P.V issue loop {np.mean(d[lo:hi + 1, 1, 2] - d[lo:hi + 1, 1, 1]):6.0f}; S issue loop {np.mean(d[lo:hi + 1, 1, 0] - d[lo:hi + 1, 1, 3]):6.0f}")
        print(f"tile 0: last pfull arrive -> P.V issued {lagpv:6.0f};  last sfree arrive -> next S issued {lags:6.0f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nc.op_attention_ex(qkv, out, batch, tokens, heads, kernel=kernel)
    e1.record(); torch.cuda.synchronize()
    print("us per launch", e0.elapsed_time(e1) / 20 * 1e3)
