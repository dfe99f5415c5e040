"""Determinism stress: every net kind / precision / batch regime of the library forwarded again and again, with idle gaps, batch
sizes interleaved (so CUDA-graph replays, programmatic dependent launches, chunk ramps and the streaming kernels all alternate),
each result compared BIT FOR BIT with the first result of the same (net, batch).  The kernels have no floating-point atomics and a
fixed reduction order, so any difference is a race (this is how the narrow-layer grid-barrier bug of mlp_umma_stream.cu shows up
without an oracle).  Diagnostics, not a test.   usage: determinism_stress.py [seconds]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(5)


def vit_case(name, cfg, precision, batches, max_batch):
    net = nc.Net.vit(cfg, max_batch=max_batch, precision=precision)
    net.upload_vit(nc.vit_random_params(cfg, seed=3))
    n_in = 3 * cfg["image_size"] ** 2
    xs = {b: rng.uniform(-1, 1, (b, n_in)).astype(np.float32) for b in batches}
    return name, net, xs, lambda n, x: n.forward(x)


def mlp_case(name, npl, n_ins, precision, batches, max_batch=0, act=0):
    net = nc.Net.mlp(npl, n_ins, precision=precision, activation=act, max_batch=max_batch)
    n_w = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    if precision == nc.PREC_INT8:
        net.upload_mlp_i8(np.clip(np.rint(rng.standard_normal(n_w) * 128.0 / np.sqrt(n_ins) * 1.4), -128, 127).astype(np.int8),
                          rng.integers(-2000, 2000, sum(npl), dtype=np.int32))
        xs = {b: rng.integers(-128, 128, (b, n_ins), dtype=np.int8) for b in batches}
        return name, net, xs, lambda n, x: n.forward_i8(x)
    net.upload_mlp((rng.standard_normal(n_w) / np.sqrt(n_ins)).astype(np.float32), rng.uniform(-0.1, 0.1, sum(npl)).astype(np.float32))
    xs = {b: rng.uniform(-1, 1, (b, n_ins)).astype(np.float32) for b in batches}
    return name, net, xs, lambda n, x: n.forward(x)


tiny2 = dict(image_size=224, patch_size=16, dim=192, depth=2, heads=3, mlp_dim=768, n_classes=1000)
base2 = dict(image_size=224, patch_size=16, dim=768, depth=2, heads=12, mlp_dim=3072, n_classes=1000)
long1 = dict(image_size=384, patch_size=16, dim=256, depth=1, heads=4, mlp_dim=512, n_classes=10)
odd = dict(image_size=96, patch_size=16, dim=128, depth=2, heads=2, mlp_dim=256, n_classes=7)
cases = [
    vit_case("vit_tiny_d2 bf16", tiny2, nc.PREC_BF16, (1, 3, 40, 96), 64),
    vit_case("vit_tiny_d2 tf32", tiny2, nc.PREC_TF32, (1, 17, 64), 64),
    vit_case("vit_base_d2 bf16", base2, nc.PREC_BF16, (1, 8, 33, 150), 64),
    vit_case("vit_577tok bf16", long1, nc.PREC_BF16, (1, 5, 20), 16),
    vit_case("vit_37tok bf16", odd, nc.PREC_BF16, (1, 7, 100, 300), 128),
    mlp_case("mlp C1 fp32", [128, 64, 10], 784, nc.PREC_FP32, (1, 64, 200)),
    mlp_case("mlp C1 tf32", [128, 64, 10], 784, nc.PREC_TF32, (1, 64, 200)),
    mlp_case("mlp C1 bf16", [128, 64, 10], 784, nc.PREC_BF16, (1, 64, 200)),
    mlp_case("mlp 6x512 bf16 (graphs)", [512] * 6, 512, nc.PREC_BF16, (1, 9, 64, 700)),
    mlp_case("mlp int8 ragged", [272, 48, 10], 1040, nc.PREC_INT8, (1, 8, 32, 33, 63, 64, 100, 128, 129, 300), 320),
    mlp_case("mlp int8 4x2048", [2048] * 4, 2048, nc.PREC_INT8, (1, 16, 32, 48, 64, 128, 129, 1024), 1024),
    mlp_case("mlp int8 wide-narrow-wide", [4096, 304, 4096], 4096, nc.PREC_INT8, (2, 31, 64, 127, 200), 256),
]
first, bad, calls = {}, {}, 0
t_end = time.time() + seconds
it = 0
while time.time() < t_end:
    for ci, (name, net, xs, fwd) in enumerate(cases):
        order = list(xs) if (it + ci) % 2 == 0 else list(xs)[::-1]
        for b in order:
            if (it + b) % 5 == 0: time.sleep(0.003)  # an idle gap: the next launch meets a GPU whose clocks have dropped
            got = fwd(net, xs[b])
            calls += 1
            key = (ci, b)
            if key not in first:
                first[key] = got.copy()
                assert np.isfinite(got.astype(np.float64)).all(), (name, b)
            elif not np.array_equal(got, first[key]):
                d = got != first[key]
                rows = np.unique(np.nonzero(d)[0])
                bad.setdefault((name, b), []).append((it, int(d.sum()), rows[:8].tolist()))
    it += 1
print(f"{it} sweeps, {calls} calls over {len(first)} (net, batch) pairs")
for k, v in bad.items():
    print("NONDETERMINISTIC", k, len(v), "times; first:", v[0])
print("all bit-identical" if not bad else f"{len(bad)} (net, batch) pairs differed")
for _, net, _, _ in cases: net.close()
