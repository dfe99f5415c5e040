"""Folded-LayerNorm path vs the float64 rounding model and the default path on small nets (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import netcuda as nc
from bf16_pipeline_model import vit_forward_bf16_model
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for (img, dim, depth, heads, mlp, batch) in ((64, 128, 1, 2, 256, 9), (64, 128, 2, 2, 256, 9), (64, 128, 2, 2, 256, 16), (64, 256, 2, 4, 512, 9), (32, 192, 3, 3, 384, 32), (64, 128, 2, 2, 256, 30)):
    cfg = dict(image_size=img, patch_size=16, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, n_classes=10)
    flat = nc.vit_random_params(cfg, seed=21)
    x = np.random.default_rng(22).uniform(-1, 1, (batch, 3 * img * img)).astype(np.float32)
    net = nc.Net.vit(cfg, max_batch=32)
    net.upload_vit(flat)
    out = {}
    for fused in (False, True):
        net.set_ln_fusion(fused)
        out[fused] = net.forward(x)
    net.close()
    m = {f: vit_forward_bf16_model(cfg, flat, x.reshape(batch, 3, img, img), ln_fused=f) for f in (False, True)}
    m64 = vit_forward_bf16_model(cfg, flat, x.reshape(batch, 3, img, img), rounding=False)
    print(cfg, "batch", batch)
    print("  unfused vs model %.2e | fused vs fused-model %.2e | fused vs unfused %.2e | models apart %.2e | unfused vs f64 %.2e | fused vs f64 %.2e" % (
        rel(out[False], m[False]), rel(out[True], m[True]), rel(out[True], out[False]), rel(m[True], m[False]), rel(out[False], m64), rel(out[True], m64)))
    per = np.abs(out[True] - m[True]).max(1) / np.abs(m[True]).max()
    print("  per-image err fused vs model:", np.array2string(per, precision=1))
