import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import netcuda as nc
from bf16_pipeline_model import vit_forward_bf16_model
g = np.load(os.path.join(ROOT, "tests", "golden", "vit_small.npz"))
cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), [int(v) for v in g["cfg"]]))
model = vit_forward_bf16_model(cfg, g["flat"], g["images"])
rel = lambda a, r: float((np.abs(a - r).max(1) / np.abs(r).max(1)).max())
for variant in (0, 1, 2):
    net = nc.Net.vit(cfg); net.upload_vit(g["flat"]); net.set_gemm_variant(variant)
    outs = [net.forward(g["images"].reshape(4, -1)) for _ in range(3)]
    net.close()
    print("variant", variant, "vs model", [rel(o, model) for o in outs], "vs fp32", rel(outs[0], g["logits"]), flush=True)
