"""clock64 timeline (cluster 0) of the last MODE_LNFOLD / MODE_RESLN GEMM of a one-block ViT-B pass (fc1 / proj)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(40 * 24 * 4, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_GEMM_DEBUG_PTR"] = hex(dbg.data_ptr())
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
os.environ["NETCUDA_GEMM_DEBUG_MODE"] = str(mode)
import netcuda as nc
np.set_printoptions(linewidth=250)
cfg = dict(image_size=224, patch_size=16, dim=768, depth=1, heads=12, mlp_dim=3072, n_classes=1000)
net = nc.Net.vit(cfg, max_batch=256)
net.upload_vit(nc.vit_random_params(cfg, seed=1))
net.set_ln_fusion(mode != 0)
x = np.random.default_rng(0).uniform(-1, 1, (256, 3 * 224 * 224)).astype(np.float32)
for _ in range(3):
    dbg.zero_()
    net.forward(x)
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(40, 24, 4)
t0 = d[0, 11, 0]
print("==== mode", mode)
print("MMA issuer: wait tempty start, tempty ready, last commit issued  (per tile, rel cycles)")
print((d[:20, 11, :3] - t0).T)
for wv in (0, 4, 12 + 7):
    print(f"epilogue warp {wv % 12} of CTA {wv // 12}: tile top, wait tfull start, tfull ready, reads done")
    print((d[:20, wv, [3, 0, 1, 2]] - t0).T)
