"""cuBLASLt int8 GEMM rate on this GPU (torch._int_mm) at the C5 layer shape, next to the netCUDA layer -- the denominator for C5's tensor-bound regime."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for m in (4096, 16384):
    a = torch.randint(-128, 128, (m, 4096), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 128, (4096, 4096), dtype=torch.int8, device="cuda").t().contiguous().t()  # column-major B, as cuBLASLt wants it
    ms = timeit(lambda: torch._int_mm(a, b))
    print(f"cuBLASLt int8 {m}x4096x4096: {ms * 1e3:.1f} us -> {2.0 * m * 4096 * 4096 / ms / 1e9:.0f} TOP/s")
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=16384); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for m in (4096, 16384):
    x = torch.randint(-128, 128, (m, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((m, 4096), dtype=torch.int32, device="cuda")
    for _ in range(3): net.forward_device_i8(x, y, m, s)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): net.forward_device_i8(x, y, m, s)
    e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"netCUDA 8-layer int8 MLP batch {m}: {ms * 1e3:.1f} us -> {2.0 * 8 * m * 4096 * 4096 / ms / 1e9:.0f} TOP/s (requantising epilogues included)")
