"""CPU tests: pin the oracle (oracle/*.c) against every fixed point that exists for this path.

The reference ships no tests and no golden vectors, so the pins are:
  * the glibc-rand() random-init rule of src/netFPGA.cpp:82-88 as a known-answer test;
  * tests/golden/mlp_c1.npz: outputs of the reference's own host runtime (unmodified src/netFPGA.cpp
    over the OpenCL shim) -- and, when oracle/_ref was built in this container, that runtime live;
  * tests/golden/vit_small.npz and a live torchvision VisionTransformer for the ViT restatement.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, c1_net, rel_err


def test_rand_init_known_answer(oracle):
    # first values quoted in SURVEY.md s.8c for srand(1) (glibc default seed)
    w, b = oracle.rand_init(1, 109184, 202)
    np.testing.assert_array_equal(w[:8], np.array([0.83, -0.14, 0.77, 0.15, 0.93, 0.35, 0.86, -0.08], dtype=np.float32))
    g = np.load(os.path.join(GOLDEN, "rand_kat.npz"))
    np.testing.assert_array_equal(w[:16], g["w16"])
    np.testing.assert_array_equal(b[:16], g["b16"])
    assert float(w.astype(np.float64).sum()) == float(g["w_sum"])
    assert w.min() >= -1.0 and w.max() <= 0.99


def test_mlp_batched_equals_scalar_statement(oracle):
    rng = np.random.default_rng(7)
    for npl, n_ins in ([5, 3], 7), ([128, 64, 10], 784), ([33, 17, 9, 4], 61):
        n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
        w = rng.uniform(-1, 1, n_params).astype(np.float32)
        b = rng.uniform(-1, 1, sum(npl)).astype(np.float32)
        x = rng.uniform(-1, 1, (9, n_ins)).astype(np.float32)
        for act in (0, 1, 2):
            batched = oracle.mlp_forward(x, w, b, npl, n_ins, act)
            scalar = np.stack([oracle.mlp_forward_one(x[i], w, b, npl, n_ins, act) for i in range(len(x))])
            np.testing.assert_array_equal(batched, scalar)


def test_mlp_hand_computed(oracle):
    # 2-3-2 net small enough to verify by hand: layer0 relu, layer1 linear
    w = np.array([1, 2, -1, 0.5, 0, -3, 1, 1, 1, -1, 0, 2], dtype=np.float32)
    b = np.array([0.5, -0.25, 1, 0, 1], dtype=np.float32)
    x = np.array([[1.0, -2.0]], dtype=np.float32)
    h = np.maximum(np.array([1 - 4 + 0.5, -1 - 1 - 0.25, 0 + 6 + 1], dtype=np.float32), 0)  # [0, 0, 7]
    want = np.array([[h.sum() + 0, -h[0] + 2 * h[2] + 1]], dtype=np.float32)  # [7, 15]
    np.testing.assert_array_equal(oracle.mlp_forward(x, w, b, [3, 2], 2), want)
    # RELU_ALL clamps the output layer too; NONE leaves hidden negatives alive
    np.testing.assert_array_equal(oracle.mlp_forward(-x, w, b, [3, 2], 2, act=1) >= 0, True)
    lin = oracle.mlp_forward(x, w, b, [3, 2], 2, act=2)
    hl = np.array([-2.5, -2.25, 7], dtype=np.float32)
    np.testing.assert_allclose(lin, [[hl.sum(), -hl[0] + 2 * hl[2] + 1]], rtol=1e-6)


def test_golden_c1_reference_runtime(oracle):
    g = np.load(os.path.join(GOLDEN, "mlp_c1.npz"))
    npl, n_ins, w, b = c1_net(oracle)
    assert list(g["npl"]) == npl and int(g["n_ins"]) == n_ins
    np.testing.assert_array_equal(oracle.mlp_forward(g["x"], w, b, npl, n_ins), g["y"])


def test_live_reference_runtime_matches_oracle(oracle):
    from oracle import Reference

    if not Reference.available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = Reference()
    rng = np.random.default_rng(3)
    for npl, n_ins, seed in ([16, 8, 4], 12, 5), ([128, 64, 10], 784, 1):
        h = ref.create(npl, n_ins, random=True, seed=seed)
        w, b, n_out = ref.flat(h)
        w2, b2 = oracle.rand_init(seed, w.size, b.size)
        np.testing.assert_array_equal(w, w2)  # same rule, same order: params first, then biases
        np.testing.assert_array_equal(b, b2)
        x = rng.uniform(-1, 1, (6, n_ins)).astype(np.float32)
        np.testing.assert_array_equal(ref.forward(h, x, n_ins, n_out), oracle.mlp_forward(x, w, b, npl, n_ins))
        assert ref.forward_us(h) >= 0
        ref.destroy(h)
    # explicit weights take the data.params[layer][neuron][k] path of the constructor (src/netFPGA.cpp:89-107)
    npl, n_ins = [7, 5], 9
    w = rng.uniform(-1, 1, 9 * 7 + 7 * 5).astype(np.float32)
    b = rng.uniform(-1, 1, 12).astype(np.float32)
    h = ref.create(npl, n_ins, w, b)
    w3, b3, _ = ref.flat(h)
    np.testing.assert_array_equal(w3, w)
    np.testing.assert_array_equal(b3, b)
    ref.destroy(h)


def test_int8_oracle_properties(oracle):
    rng = np.random.default_rng(11)
    npl, n_ins = [64, 64, 32], 48
    n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = rng.integers(-128, 128, n_params, dtype=np.int8)
    bq = rng.integers(-(1 << 14), 1 << 14, sum(npl), dtype=np.int32)
    xq = rng.integers(-128, 128, (5, n_ins), dtype=np.int8)
    got = oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins)
    # independent numpy statement of the Q1.7 rule
    h = xq.astype(np.int64)
    off_w = off_b = 0
    fan_in = n_ins
    for li, fo in enumerate(npl):
        W = wq[off_w:off_w + fo * fan_in].reshape(fo, fan_in).astype(np.int64)
        acc = h @ W.T + bq[off_b:off_b + fo].astype(np.int64)
        if li < len(npl) - 1:
            h = np.minimum(127, np.maximum(acc, 0) >> 7)
        else:
            h = acc
        off_w += fo * fan_in
        off_b += fo
        fan_in = fo
    np.testing.assert_array_equal(got, h.astype(np.int32))
    # thread count must not matter (integer arithmetic, per-sample independence)
    np.testing.assert_array_equal(got, oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins, threads=1))
    # quantisers
    np.testing.assert_array_equal(oracle.quantize_q17(np.array([0.0, 1.0, -1.0, 0.5, 0.00390625, -0.00390625, 0.99], np.float32)),
                                  np.array([0, 127, -128, 64, 0, 0, 127], dtype=np.int8))
    np.testing.assert_array_equal(oracle.quantize_bias(np.array([1.0, -0.5, 1e-5], np.float32)), np.array([16384, -8192, 0], np.int32))


def test_vit_golden_fixture(oracle):
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), [int(v) for v in g["cfg"]]))
    got = oracle.vit_forward(cfg, g["flat"], g["images"])
    assert np.abs(got - g["logits"]).max() <= 1e-4
    np.testing.assert_array_equal(got.argmax(1), g["logits"].argmax(1))
    # batch order / thread count independence
    np.testing.assert_array_equal(oracle.vit_forward(cfg, g["flat"], g["images"][::-1].copy(), threads=1)[::-1], got)


def test_vit_matches_live_torchvision(oracle):
    torch = pytest.importorskip("torch")
    pytest.importorskip("torchvision")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    # ViT-Tiny width/heads at reduced depth and resolution: 3 heads, 37 tokens (not a tile multiple)
    cfg = dict(image_size=96, patch_size=16, dim=192, depth=3, heads=3, mlp_dim=768, n_classes=50)
    model = mg.make_torchvision_vit(cfg, seed=3)
    flat = mg.flatten_torchvision_vit(model)
    assert flat.size == oracle.vit_param_count(cfg)
    x = np.random.default_rng(5).uniform(-1, 1, (3, 3, 96, 96)).astype(np.float32)
    with torch.no_grad():
        want = model(torch.from_numpy(x)).numpy()
    got = oracle.vit_forward(cfg, flat, x)
    assert np.abs(got - want).max() <= 1e-4, np.abs(got - want).max()


def test_building_blocks(oracle):
    rng = np.random.default_rng(2)
    a = rng.standard_normal((37, 50)).astype(np.float32)
    w = rng.standard_normal((23, 50)).astype(np.float32)
    bias = rng.standard_normal(23).astype(np.float32)
    np.testing.assert_allclose(oracle.linear(a, w, bias), a.astype(np.float64) @ w.T.astype(np.float64) + bias, rtol=2e-5, atol=2e-5)
    x = rng.standard_normal((9, 64)).astype(np.float32) * 3 + 1
    g, b = rng.standard_normal(64).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    xd = x.astype(np.float64)
    want = (xd - xd.mean(1, keepdims=True)) / np.sqrt(xd.var(1, keepdims=True) + 1e-6) * g + b
    np.testing.assert_allclose(oracle.layernorm(x, g, b), want, rtol=1e-5, atol=1e-5)
    # attention vs a dense numpy softmax, 2 images x 3 heads x 10 tokens
    B, T, H = 2, 10, 3
    qkv = rng.standard_normal((B * T, 3 * H * 64)).astype(np.float32)
    got = oracle.attention(qkv, B, T, H)
    q3 = qkv.reshape(B, T, 3, H, 64).astype(np.float64)
    for bi in range(B):
        for h in range(H):
            s = q3[bi, :, 0, h] @ q3[bi, :, 1, h].T / 8.0
            p = np.exp(s - s.max(1, keepdims=True))
            p /= p.sum(1, keepdims=True)
            np.testing.assert_allclose(got.reshape(B, T, H, 64)[bi, :, h], p @ q3[bi, :, 2, h], rtol=1e-4, atol=1e-5)
    from math import erf, sqrt
    xs = np.linspace(-5, 5, 41).astype(np.float32)
    np.testing.assert_allclose(oracle.gelu(xs), [0.5 * v * (1 + erf(v / sqrt(2))) for v in xs.astype(np.float64)], rtol=1e-5, atol=1e-6)
