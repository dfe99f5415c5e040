"""Dump the GPU logits of the golden ViT fixture (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
g = np.load(os.path.join(ROOT, "tests", "golden", "vit_small.npz"))
cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), [int(v) for v in g["cfg"]]))
out = {}
for variant in (0, 1):
    net = nc.Net.vit(cfg)
    net.upload_vit(g["flat"])
    net.set_gemm_variant(variant)
    out[f"v{variant}"] = net.forward(g["images"].reshape(4, -1))
    net.close()
np.savez(os.path.join(ROOT, "gpurun_out", "vit_small_gpu.npz"), **out)
print("dumped")
