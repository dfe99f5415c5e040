"""Generate the committed golden fixtures under tests/golden/.

Run once in the CPU container:  python tests/golden/make_golden.py

Fixtures (all small .npz files):
  vit_small.npz   a 2-block ViT (image 32, patch 16, D=128, 2 heads, mlp 256, 10 classes) built with
                  torchvision.models.vision_transformer.VisionTransformer under torch.manual_seed(0):
                  flat weights in the netcuda/oracle layout, 4 input images, torchvision's logits.
                  Pins oracle_vit_forward to an independent implementation of the published algorithm
                  (the reference has no ViT to compare against).
  mlp_c1.npz      config C1 (784-128-64-10, batch 64): weights from the reference's random-init rule
                  with srand(1) (src/netFPGA.cpp:82-88), inputs uniform[-1,1) seed 1234, outputs of the
                  reference's own host runtime (oracle/_ref: unmodified src/netFPGA.cpp over the OpenCL
                  shim, kernel = oracle_mlp_forward_one).  Pins layout + I/O contract.
  rand_kat.npz    first 16 params / biases of that random init (known-answer for the glibc rand() rule).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def flatten_torchvision_vit(model) -> np.ndarray:
    """torchvision VisionTransformer -> flat fp32 vector in the layout of netcuda_vit_param_count."""
    sd = {k: v.detach().cpu().float().numpy() for k, v in model.state_dict().items()}
    parts = [sd["conv_proj.weight"].reshape(sd["conv_proj.weight"].shape[0], -1), sd["conv_proj.bias"],
             sd["class_token"].reshape(-1), sd["encoder.pos_embedding"].reshape(-1)]
    depth = len(model.encoder.layers)
    for i in range(depth):
        p = f"encoder.layers.encoder_layer_{i}."
        parts += [sd[p + "ln_1.weight"], sd[p + "ln_1.bias"],
                  sd[p + "self_attention.in_proj_weight"], sd[p + "self_attention.in_proj_bias"],
                  sd[p + "self_attention.out_proj.weight"], sd[p + "self_attention.out_proj.bias"],
                  sd[p + "ln_2.weight"], sd[p + "ln_2.bias"],
                  sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"], sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"]]
    parts += [sd["encoder.ln.weight"], sd["encoder.ln.bias"], sd["heads.head.weight"], sd["heads.head.bias"]]
    return np.concatenate([np.ascontiguousarray(p, dtype=np.float32).ravel() for p in parts])


def make_torchvision_vit(cfg: dict, seed: int = 0):
    from torchvision.models.vision_transformer import VisionTransformer

    torch.manual_seed(seed)
    m = VisionTransformer(image_size=cfg["image_size"], patch_size=cfg["patch_size"], num_layers=cfg["depth"],
                          num_heads=cfg["heads"], hidden_dim=cfg["dim"], mlp_dim=cfg["mlp_dim"],
                          num_classes=cfg["n_classes"])
    # torchvision zero-initialises the head and the biases; give every tensor a non-trivial value so
    # that the fixture exercises all of them.
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        m.heads.head.weight.copy_(torch.randn(m.heads.head.weight.shape, generator=g) * 0.05)
        for name, p in m.named_parameters():
            if name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            if "ln" in name and name.endswith("weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.05)
        m.class_token.copy_(torch.randn(m.class_token.shape, generator=g) * 0.02)
    return m.eval()


def main():
    from oracle import Oracle, Reference

    cfg = dict(image_size=32, patch_size=16, dim=128, depth=2, heads=2, mlp_dim=256, n_classes=10)
    model = make_torchvision_vit(cfg)
    flat = flatten_torchvision_vit(model)
    rng = np.random.default_rng(1234)
    images = rng.uniform(-1, 1, (4, 3, 32, 32)).astype(np.float32)
    with torch.no_grad():
        logits = model(torch.from_numpy(images)).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "vit_small.npz"), flat=flat, images=images, logits=logits,
                        cfg=np.array([cfg[k] for k in ("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim",
                                                       "n_classes")], dtype=np.int32))
    o = Oracle()
    mine = o.vit_forward(cfg, flat, images)
    print("vit_small: max|oracle - torchvision| =", float(np.abs(mine - logits).max()), "max|logit| =",
          float(np.abs(logits).max()))

    npl, n_ins = [128, 64, 10], 784
    n_params = 784 * 128 + 128 * 64 + 64 * 10
    w, b = o.rand_init(1, n_params, sum(npl))
    x = np.random.default_rng(1234).uniform(-1, 1, (64, n_ins)).astype(np.float32)
    if Reference.available():
        r = Reference()
        h = r.create(npl, n_ins, random=True, seed=1)
        w_ref, b_ref, n_out = r.flat(h)
        assert np.array_equal(w_ref, w) and np.array_equal(b_ref, b)
        y = r.forward(h, x, n_ins, n_out)
        r.destroy(h)
        np.savez_compressed(os.path.join(HERE, "mlp_c1.npz"), x=x, y=y, seed=np.int32(1),
                            npl=np.array(npl, dtype=np.int32), n_ins=np.int32(n_ins))
        print("mlp_c1: reference outputs saved; equal to oracle:", np.array_equal(y, o.mlp_forward(x, w, b, npl, n_ins)))
    np.savez_compressed(os.path.join(HERE, "rand_kat.npz"), w16=w[:16], b16=b[:16], w_sum=np.float64(w.astype(np.float64).sum()),
                        b_sum=np.float64(b.astype(np.float64).sum()))


if __name__ == "__main__":
    main()
