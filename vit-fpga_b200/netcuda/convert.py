"""Checkpoint import: state dicts of the two common ViT implementations -> the flat fp32 vector of
netcuda_vit_param_count (include/netcuda.h), ready for netcuda_upload_vit / netcuda_file_write_vit.

The reference has no model zoo; its only notion of a stored net is the flat W[out][in] layout of
src/netFPGA.cpp:91-106, which is what these functions produce.  Pure numpy: values may be
torch tensors or arrays.  Linear weights are already [out][in] in both sources, so nothing is transposed.
"""
from __future__ import annotations

import numpy as np


def _np(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().float().numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


def cfg_from_torchvision(sd: dict) -> dict:
    w = _np(sd["conv_proj.weight"])
    dim, _, p, _ = w.shape
    tokens = _np(sd["encoder.pos_embedding"]).shape[1]
    depth = 1 + max(int(k.split("encoder_layer_")[1].split(".")[0]) for k in sd if "encoder_layer_" in k)
    grid = int(round((tokens - 1) ** 0.5))
    return dict(image_size=grid * p, patch_size=p, dim=dim, depth=depth, heads=dim // 64,
                mlp_dim=_np(sd["encoder.layers.encoder_layer_0.mlp.0.weight"]).shape[0],
                n_classes=_np(sd["heads.head.weight"]).shape[0])


def flat_from_torchvision(sd: dict) -> np.ndarray:
    """torchvision.models.vision_transformer.VisionTransformer.state_dict() -> flat vector."""
    cfg = cfg_from_torchvision(sd)
    parts = [_np(sd["conv_proj.weight"]).reshape(cfg["dim"], -1), _np(sd["conv_proj.bias"]), _np(sd["class_token"]).reshape(-1),
             _np(sd["encoder.pos_embedding"]).reshape(-1)]
    for i in range(cfg["depth"]):
        p = f"encoder.layers.encoder_layer_{i}."
        parts += [_np(sd[p + k]) for k in ("ln_1.weight", "ln_1.bias", "self_attention.in_proj_weight", "self_attention.in_proj_bias",
                                          "self_attention.out_proj.weight", "self_attention.out_proj.bias", "ln_2.weight", "ln_2.bias",
                                          "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias")]
    parts += [_np(sd[k]) for k in ("encoder.ln.weight", "encoder.ln.bias", "heads.head.weight", "heads.head.bias")]
    return np.concatenate([p.ravel() for p in parts])


def cfg_from_hf(sd: dict, prefix: str = "vit.") -> dict:
    w = _np(sd[prefix + "embeddings.patch_embeddings.projection.weight"])
    dim, _, p, _ = w.shape
    tokens = _np(sd[prefix + "embeddings.position_embeddings"]).shape[1]
    depth = 1 + max(int(k.split("encoder.layer.")[1].split(".")[0]) for k in sd if "encoder.layer." in k)
    grid = int(round((tokens - 1) ** 0.5))
    return dict(image_size=grid * p, patch_size=p, dim=dim, depth=depth, heads=dim // 64,
                mlp_dim=_np(sd[prefix + "encoder.layer.0.intermediate.dense.weight"]).shape[0],
                n_classes=_np(sd["classifier.weight"]).shape[0])


def flat_from_hf(sd: dict, prefix: str = "vit.") -> np.ndarray:
    """transformers.ViTForImageClassification.state_dict() -> flat vector (q, k, v matrices are stacked into the
    single [3D][D] in-projection, q|k|v order, which is the layout the attention kernel reads)."""
    cfg = cfg_from_hf(sd, prefix)
    e = prefix + "embeddings."
    parts = [_np(sd[e + "patch_embeddings.projection.weight"]).reshape(cfg["dim"], -1), _np(sd[e + "patch_embeddings.projection.bias"]),
             _np(sd[e + "cls_token"]).reshape(-1), _np(sd[e + "position_embeddings"]).reshape(-1)]
    for i in range(cfg["depth"]):
        p = f"{prefix}encoder.layer.{i}."
        a = p + "attention.attention."
        parts += [_np(sd[p + "layernorm_before.weight"]), _np(sd[p + "layernorm_before.bias"]),
                  np.concatenate([_np(sd[a + n + ".weight"]) for n in ("query", "key", "value")], axis=0),
                  np.concatenate([_np(sd[a + n + ".bias"]) for n in ("query", "key", "value")], axis=0),
                  _np(sd[p + "attention.output.dense.weight"]), _np(sd[p + "attention.output.dense.bias"]),
                  _np(sd[p + "layernorm_after.weight"]), _np(sd[p + "layernorm_after.bias"]),
                  _np(sd[p + "intermediate.dense.weight"]), _np(sd[p + "intermediate.dense.bias"]),
                  _np(sd[p + "output.dense.weight"]), _np(sd[p + "output.dense.bias"])]
    parts += [_np(sd[prefix + "layernorm.weight"]), _np(sd[prefix + "layernorm.bias"]), _np(sd["classifier.weight"]), _np(sd["classifier.bias"])]
    return np.concatenate([p.ravel() for p in parts])
