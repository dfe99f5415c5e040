"""Torch-CPU model of the BF16 ViT pipeline's rounding points (test infrastructure).

The CUDA path keeps the residual stream, biases, LayerNorm statistics, softmax statistics and every
accumulator in fp32 and rounds to bf16 exactly where a tensor-core operand is produced: weights, patch
pixels, LayerNorm outputs, the packed qkv, the un-normalised softmax numerators P, the attention output
and the GELU output.  This model applies the same roundings to an otherwise fp32 forward, so that

    |cuda - model|   measures kernel arithmetic (accumulation order, exp2/erf approximations), and
    |model - fp32|   is the budget that bf16 operands cost -- not something a kernel can win back.

Everything between two rounding points is evaluated in float64, so the model does not depend on which
matmul kernels (and which reduced-precision fast paths) the host CPU's BLAS picks.
"""
import numpy as np
import torch


def _bf(x):
    return x.float().to(torch.bfloat16).double()


def vit_forward_bf16_model(cfg: dict, flat: np.ndarray, images: np.ndarray, rounding: bool = True) -> np.ndarray:
    r = _bf if rounding else (lambda t: t)
    flat = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64))
    D, F, C, P, S, H = cfg["dim"], cfg["mlp_dim"], cfg["n_classes"], cfg["patch_size"], cfg["image_size"], cfg["heads"]
    g = S // P
    NP, PK = g * g, 3 * P * P
    T = NP + 1
    imgs = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float64)).reshape(-1, 3, S, S)
    B = imgs.shape[0]
    pos_ = [0]

    def take(*shape):
        n = int(np.prod(shape))
        v = flat[pos_[0]:pos_[0] + n].reshape(*shape)
        pos_[0] += n
        return v

    pw, pb, cls, pos = take(D, PK), take(D), take(D), take(T, D)
    patches = imgs.reshape(B, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B, NP, PK)
    x = torch.empty(B, T, D, dtype=torch.float64)
    x[:, 1:] = r(patches) @ r(pw).T + pb + pos[1:]
    x[:, 0] = cls + pos[0]
    ln = torch.nn.functional.layer_norm
    for _ in range(cfg["depth"]):
        g1, b1, qw, qb, ow, ob = take(D), take(D), take(3 * D, D), take(3 * D), take(D, D), take(D)
        g2, b2, f1w, f1b, f2w, f2b = take(D), take(D), take(F, D), take(F), take(D, F), take(D)
        y = r(ln(x, (D,), g1, b1, 1e-6))
        qkv = r(y @ r(qw).T + qb).reshape(B, T, 3, H, D // H)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        s = (q @ k.transpose(-1, -2)) / float(np.sqrt(D // H))
        pe = torch.exp(s - s.max(-1, keepdim=True).values)
        # the row sum comes out of the tensor core next to P.V (P.1 against a tile of ones, csrc/attention.cu): it is the sum of the
        # bf16-rounded numerators, the very weights that multiply V
        o = (r(pe) @ v) / r(pe).sum(-1, keepdim=True)
        x = x + r(o.permute(0, 2, 1, 3).reshape(B, T, D)) @ r(ow).T + ob
        y = r(ln(x, (D,), g2, b2, 1e-6))
        h = r(torch.nn.functional.gelu(y @ r(f1w).T + f1b))
        x = x + h @ r(f2w).T + f2b
    gf, bfin, hw, hb = take(D), take(D), take(C, D), take(C)
    assert pos_[0] == flat.numel()
    c = r(ln(x[:, 0], (D,), gf, bfin, 1e-6))
    return (c @ r(hw).T + hb).float().numpy()
