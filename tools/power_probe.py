"""Where do the joules of a ViT-B step go?  The B200 sits at its 1 kW software power cap during the bench, so the step time is
energy / 1 kW.  Every kernel class is run back to back for ~1.5 s while nvidia-smi samples power and SM clock;
energy per launch = median power x time per launch.  torch.matmul (cuBLAS) at the same shapes is the yardstick for the GEMMs."""
import os, subprocess, sys, threading, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc

rows_lock = threading.Lock(); rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
def reader():
    for line in proc.stdout:
        try:
            c, p = [float(v) for v in line.split(",")]
        except ValueError:
            continue
        with rows_lock: rows.append((time.perf_counter(), c, p))
threading.Thread(target=reader, daemon=True).start()

B, T, D, F, H = 512, 197, 768, 3072, 12
M = B * T
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
bf = torch.bfloat16
y = torch.randn((M, D), device="cuda").to(bf)
hid = torch.randn((M, F), device="cuda").to(bf)
x = torch.randn((M, D), device="cuda")
qkv = torch.randn((M, 3 * D), device="cuda").to(bf)
att = torch.empty((M, D), dtype=bf, device="cuda")
w_qkv = (torch.randn((3 * D, D), device="cuda") * 0.03).to(bf); w_fc1 = (torch.randn((F, D), device="cuda") * 0.03).to(bf)
w_fc2 = (torch.randn((D, F), device="cuda") * 0.02).to(bf); w_proj = (torch.randn((D, D), device="cuda") * 0.03).to(bf)
b3 = torch.randn(3 * D, device="cuda"); bF = torch.randn(F, device="cuda"); bD = torch.randn(D, device="cuda")
g = torch.ones(D, device="cuda"); be = torch.zeros(D, device="cuda")
o_f32 = torch.empty((M, F), device="cuda")
ops = {
    "qkv": (lambda: nc.op_gemm(y, w_qkv, b3, qkv, nc.PREC_BF16, nc.OUT_BF16, epilogue=nc.EPI_NONE, stream=s), 2.0 * M * 3 * D * D),
    "fc1_gelu": (lambda: nc.op_gemm(y, w_fc1, bF, hid, nc.PREC_BF16, nc.OUT_BF16, epilogue=nc.EPI_GELU, stream=s), 2.0 * M * F * D),
    "fc1_none": (lambda: nc.op_gemm(y, w_fc1, bF, hid, nc.PREC_BF16, nc.OUT_BF16, epilogue=nc.EPI_NONE, stream=s), 2.0 * M * F * D),
    "fc2_residual": (lambda: nc.op_gemm(hid, w_fc2, bD, x, nc.PREC_BF16, nc.OUT_F32, epilogue=nc.EPI_RESIDUAL, stream=s), 2.0 * M * F * D),
    "proj_residual": (lambda: nc.op_gemm(att, w_proj, bD, x, nc.PREC_BF16, nc.OUT_F32, epilogue=nc.EPI_RESIDUAL, stream=s), 2.0 * M * D * D),
    "attention": (lambda: nc.op_attention(qkv, att, B, T, H, stream=s), 4.0 * B * T * T * D),
    "layernorm": (lambda: nc.op_layernorm(x, g, be, y, stream=s), 0.0),
    "cublas_fc1_shape": (lambda: torch.matmul(y, w_fc1.t(), out=hid), 2.0 * M * F * D),
    "cublas_fc2_shape": (lambda: torch.matmul(hid, w_fc2.t(), out=att), 2.0 * M * F * D),
    "cublas_qkv_shape": (lambda: torch.matmul(y, w_qkv.t(), out=qkv), 2.0 * M * 3 * D * D),
}
if os.environ.get("PROBE_ONLY"):
    ops = {k: ops[k] for k in os.environ["PROBE_ONLY"].split(",")}
out = {}
dur = float(os.environ.get("PROBE_SECONDS", "1.5"))
for name, (fn, flops) in ops.items():
    for _ in range(3): fn()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); n = 0
    e0.record(s)
    while time.perf_counter() - t0 < dur:
        for _ in range(20): fn()
        n += 20
        if n % 200 == 0: s.synchronize()   # keep the launch queue bounded
    e1.record(s); s.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / n
    with rows_lock:
        win = [(c, p) for (t, c, p) in rows if t0 + 0.5 * dur <= t <= t1]   # second half: the power controller has settled
    clk = sorted(c for c, _ in win)[len(win) // 2] if win else None
    pw = sorted(p for _, p in win)[len(win) // 2] if win else None
    out[name] = dict(us=round(ms * 1e3, 2), tflops=round(flops / ms / 1e9, 1) if flops else None, sm_mhz=clk, watts=pw,
                     mj_per_launch=round(pw * ms, 2) if pw else None, samples=len(win))
    print(name, out[name], flush=True)
    time.sleep(0.3)
proc.terminate()
step = {"qkv": 24, "fc1_gelu": 24, "fc2_residual": 24, "proj_residual": 24, "attention": 24, "layernorm": 48}
if all(k in out and out[k]["mj_per_launch"] for k in step):
    tot = sum(out[k]["mj_per_launch"] * v for k, v in step.items()) / 1e3
    print("joules per 1024-image step (sum of kernel classes):", round(tot, 2), {k: round(out[k]["mj_per_launch"] * v / 1e3, 2) for k, v in step.items()})
    print("time per step if run back to back at these sustained rates (ms):", round(sum(out[k]["us"] * v for k, v in step.items()) / 1e3, 2))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "power_probe.json"), "w"), indent=1)
