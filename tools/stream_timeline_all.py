"""globaltimer timeline of the INT8 weight-streaming kernel (mlp_stream.cu), config C5, EVERY CTA (debug build, NETCUDA_STREAM_DEBUG_CTA=-1):
per layer the earliest / median / latest CTA at each stamp, relative to the earliest layer-0 start."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
NCTA = 160
dbg = torch.zeros(NCTA * 32 * 6, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_STREAM_DEBUG_PTR"] = hex(dbg.data_ptr())
os.environ["NETCUDA_STREAM_DEBUG_CTA"] = "-1"
import netcuda as nc
np.set_printoptions(linewidth=220)
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=64); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for batch in (int(a) for a in (sys.argv[1:] or ["1"])):
    x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
    for _ in range(5): net.forward_device_i8(x, y, batch, s)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20): net.forward_device_i8(x, y, batch, s)
    e1.record(s); s.synchronize()
    full = dbg.cpu().numpy().reshape(NCTA, 32, 6).astype(np.float64)
    d = full[:, :8]
    live = d[:, 0, 0] > 0
    d = d[live]
    ee = full[live, 24, :2]
    print(f"kernel entry min/max {(ee[:, 0].min() - d[:, 0, 0].min()) / 1000:.2f} / {(ee[:, 0].max() - d[:, 0, 0].min()) / 1000:.2f} us, exit min/max "
          f"{(ee[:, 1].min() - d[:, 0, 0].min()) / 1000:.2f} / {(ee[:, 1].max() - d[:, 0, 0].min()) / 1000:.2f} us (relative to the earliest layer-0 start)")
    t0 = d[:, 0, 0].min()
    d = (d - t0) / 1000.0
    print(f"batch {batch}: {int(live.sum())} CTAs stamped; event time per launch (20 back to back) {e0.elapsed_time(e1) * 50:.1f} us; layer-0 starts spread {d[:, 0, 0].max():.2f} us; last stamp {d[:, 7, 5].max():.2f} us")
    print("layer: [start, barrier passed, act loaded, tiles done, partials visible, released] as min / median / max over CTAs (us)")
    for l in range(8):
        print(l, " ".join(f"{np.min(d[:, l, k]):6.2f}/{np.median(d[:, l, k]):6.2f}/{np.max(d[:, l, k]):6.2f}" for k in range(6)))
    ex = (full[live, 16:24, :3] - t0) / 1000.0
    if (full[live, 16:24, 2] > 0).any():
        print("grid-barrier release, median over CTAs, us after 'partials visible': [stores issued, named barrier passed, fence done, (released)]")
        for l in range(7):
            print(l, " ".join(f"{np.median(ex[:, l, k] - d[:, l, 4]):6.2f}" for k in range(3)), f"{np.median(d[:, l, 5] - d[:, l, 4]):6.2f}")
    slow = np.argsort(-d[:, 7, 5])[:5]
    print("latest CTAs at the end:", slow.tolist(), d[slow, 7, 5].round(2).tolist())
