"""Net descriptions and synthetic weights shared by bench.py, the tests and the ctypes binding.

Pure Python + numpy: importing this module does NOT load libnetcuda.so, so bench.py's reference arm (the CPU
implementation timed alone) can use the same configurations and the same random-init weights as the GPU arm
without mapping the product library into its process.

The flat ViT parameter layout is the one documented at netcuda_vit_param_count (include/netcuda.h); the MLP
layout is the reference's flat W[out][in] / bias order (src/netFPGA.cpp:91-106).
"""
from __future__ import annotations

import numpy as np

VIT_PRESETS = {
    "vit_tiny_16_224": dict(image_size=224, patch_size=16, dim=192, depth=12, heads=3, mlp_dim=768, n_classes=1000),
    "vit_base_16_224": dict(image_size=224, patch_size=16, dim=768, depth=12, heads=12, mlp_dim=3072, n_classes=1000),
    "vit_large_16_384": dict(image_size=384, patch_size=16, dim=1024, depth=24, heads=16, mlp_dim=4096, n_classes=1000),
}

# BASELINE.json configs 1 and 5
MLP_C1 = dict(npl=[128, 64, 10], n_ins=784)
MLP_C5 = dict(npl=[4096] * 8, n_ins=4096)


def vit_param_count(cfg: dict) -> int:
    """Number of fp32 values of the flat ViT parameter vector (same formula as vit_param_count in csrc/runtime.cu)."""
    D, F, C = cfg["dim"], cfg["mlp_dim"], cfg["n_classes"]
    g = cfg["image_size"] // cfg["patch_size"]
    N, pk = g * g + 1, 3 * cfg["patch_size"] ** 2
    n = D * pk + D + D + N * D
    n += cfg["depth"] * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + F * D + F + D * F + D)
    n += 2 * D + C * D + C
    return n


def vit_flops_per_image(cfg: dict) -> float:
    """Algorithmic FLOPs (2 x MACs of the dense contractions): 2 Np 3p^2 D + L (24 N D^2 + 4 N^2 D) + 2 D C
    for an MLP ratio of 4; written out for any mlp_dim."""
    D, F, C = cfg["dim"], cfg["mlp_dim"], cfg["n_classes"]
    g = cfg["image_size"] // cfg["patch_size"]
    NP = g * g
    N, pk = NP + 1, 3 * cfg["patch_size"] ** 2
    return 2.0 * (NP * pk * D + cfg["depth"] * (N * (4.0 * D * D + 2.0 * D * F) + 2.0 * N * N * D) + D * C)


def vit_random_params(cfg: dict, seed: int = 0) -> np.ndarray:
    """Random-init weights of the named architecture in the flat layout (synthetic benchmark weights).
    Matrices ~ N(0, 0.02) like torchvision's trunc-normal init, LN gamma 1 +- 0.05, small biases."""
    rng = np.random.default_rng(seed)
    D, F, Cn = cfg["dim"], cfg["mlp_dim"], cfg["n_classes"]
    g = cfg["image_size"] // cfg["patch_size"]
    N, pk = g * g + 1, 3 * cfg["patch_size"] ** 2
    parts = []

    def mat(r, c, std=0.02):
        parts.append((rng.standard_normal((r, c), dtype=np.float32) * std).ravel())

    def vec(n, mean=0.0, std=0.02):
        parts.append(mean + rng.standard_normal(n, dtype=np.float32) * std)

    mat(D, pk, std=(1.0 / pk) ** 0.5)
    vec(D), vec(D), mat(N, D)
    for _ in range(cfg["depth"]):
        vec(D, 1.0, 0.05), vec(D)
        mat(3 * D, D, std=D ** -0.5), vec(3 * D)
        mat(D, D, std=D ** -0.5), vec(D)
        vec(D, 1.0, 0.05), vec(D)
        mat(F, D, std=D ** -0.5), vec(F)
        mat(D, F, std=F ** -0.5), vec(D)
    vec(D, 1.0, 0.05), vec(D)
    mat(Cn, D, std=0.05), vec(Cn)
    flat = np.concatenate(parts).astype(np.float32)
    assert flat.size == vit_param_count(cfg)
    return flat


def mlp_param_counts(npl, n_ins) -> tuple[int, int]:
    fan_ins = [n_ins] + list(npl[:-1])
    return sum(a * b for a, b in zip(fan_ins, npl)), sum(npl)


def mlp_reference_rule_params(npl, n_ins, seed: int = 1):
    """Weights in the value set of the reference's random initialisation, float(rand() % 200 - 100) / 100 in [-1.00, 0.99]
    (src/netFPGA.cpp:82-88), drawn from numpy's generator (the glibc sequence itself is pinned by tests/golden/rand_kat.npz)."""
    rng = np.random.default_rng(seed)
    n_w, n_b = mlp_param_counts(npl, n_ins)
    w = (rng.integers(-100, 100, n_w).astype(np.float32) / np.float32(100.0)).astype(np.float32)
    b = (rng.integers(-100, 100, n_b).astype(np.float32) / np.float32(100.0)).astype(np.float32)
    return w, b


def mlp_int8_params(npl, n_ins, seed: int = 50):
    """Q1.7 weights scaled by 1.4 / sqrt(fan_in) so that hidden activations neither saturate nor die (SURVEY.md s.8d allows either
    the raw rule or this scaling for config C5; the raw +-1 rule saturates every hidden unit after one 4096-wide layer)."""
    rng = np.random.default_rng(seed)
    n_w, n_b = mlp_param_counts(npl, n_ins)
    wq = np.clip(np.rint(rng.standard_normal(n_w, dtype=np.float32) * np.float32(128.0 / np.sqrt(n_ins) * 1.4)), -128, 127).astype(np.int8)
    bq = rng.integers(-2000, 2000, n_b, dtype=np.int32)
    return wq, bq
