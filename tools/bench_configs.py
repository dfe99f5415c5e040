"""Throughput of the other BASELINE.json configs (C1, C2, C4, C5) on one B200 -- the numbers quoted in DESIGN.md s.5.
Device-resident inputs, CUDA events on an explicit stream, median of several runs.  C4 (ViT-L/16-384, 24 blocks) is run in full."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc

def timeit(fn, stream, reps=7, inner=1):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(inner):
            fn()
        e1.record(stream)
        stream.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2]

out = {}
s = torch.cuda.Stream()
torch.cuda.set_stream(s)

# C1: 784-128-64-10, batch 64 (plumbing config: launch-latency bound)
npl, n_ins = [128, 64, 10], 784
rng = np.random.default_rng(0)
w = rng.uniform(-1, 1, 784 * 128 + 128 * 64 + 64 * 10).astype(np.float32); b = rng.uniform(-1, 1, 202).astype(np.float32)
for prec in ("fp32", "tf32", "bf16"):
    net = nc.Net.mlp(npl, n_ins, precision=nc.PRECISIONS[prec]); net.upload_mlp(w, b)
    x = torch.rand((64, n_ins), device="cuda"); y = torch.empty((64, 10), device="cuda")
    for _ in range(5): net.forward_device(x, y, 64, s)
    ms = timeit(lambda: net.forward_device(x, y, 64, s), s, inner=50)
    hx = torch.rand((64, n_ins)).pin_memory(); hy = torch.empty((64, 10)).pin_memory()
    t0 = time.perf_counter()
    for _ in range(200): net.forward_into(hx, hy)
    e2e_ms = (time.perf_counter() - t0) / 200 * 1e3
    out[f"C1 {prec}"] = dict(us_per_batch_device=ms * 1e3, samples_per_s_device=64 / ms * 1e3, us_per_batch_host_call=e2e_ms * 1e3, samples_per_s_host_call=64 / e2e_ms * 1e3)
    net.close()

# C2 / C4: ViT-Tiny/16-224 b256, ViT-L/16-384 b64 per pass
for name, batch, mb in (("vit_tiny_16_224", 256, 256), ("vit_large_16_384", 64, 32)):
    cfg = nc.VIT_PRESETS[name]
    net = nc.Net.vit(cfg, max_batch=mb); net.upload_vit(nc.vit_random_params(cfg, seed=0))
    x = torch.rand((batch, net.n_in), device="cuda") * 2 - 1; y = torch.empty((batch, 1000), device="cuda")
    for _ in range(3): net.forward_device(x, y, batch, s)
    s.synchronize()
    ms = timeit(lambda: net.forward_device(x, y, batch, s), s, reps=5, inner=3)
    net.profile_enable(True); net.forward_device(x, y, batch, s); prof = net.profile_read(); net.profile_enable(False)
    out[name] = dict(batch=batch, ms=ms, images_per_s=batch / ms * 1e3, tflops=batch * net.flops_per_sample / ms / 1e9,
                     per_kernel_ms={k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:7]})
    net.close()

# C5: 8 x 4096 INT8, batch sweep
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
for path in ("default", "gemm_only"):  # default: <= 16 samples run the persistent weight-streaming kernel (mlp_stream.cu)
    if path == "gemm_only": os.environ["NETCUDA_MLP_STREAM"] = "0"
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=16384); net.upload_mlp_i8(wq, bq)
    for batch in (1, 4, 16, 128, 1024, 4096, 16384) if path == "default" else (1, 4, 16):
        x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
        for _ in range(3): net.forward_device_i8(x, y, batch, s)
        ms = timeit(lambda: net.forward_device_i8(x, y, batch, s), s, inner=5)
        ops = 2.0 * 8 * 4096 * 4096 * batch
        out[f"C5 int8 batch {batch}" + ("" if path == "default" else " (split-K GEMM path)")] = dict(
            ms=ms, samples_per_s=batch / ms * 1e3, tops=ops / ms / 1e9, weight_stream_gbs=8 * 4096 * 4096 / ms / 1e6)
    net.close()
print(json.dumps(out, indent=1))
