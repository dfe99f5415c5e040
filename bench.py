#!/usr/bin/env python
"""bench.py -- images/sec of the batched forward pass through the netCUDA backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl netcuda|reference] [--workload NAME] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of one batch of synthetic inputs through the net.  The headline workload is BASELINE.json's
config 3: ViT-B/16-224, 1024 images per GPU per step, bf16, random-init weights.  One process per GPU; every rank forwards its
own shard of the global batch with replicated weights (no data-path collective), then the logits are gathered with one NCCL
all-gather (N > 1) on a side stream that overlaps the next step -- as BASELINE.json's north_star describes.

One JSON line on rank 0:
  value        whole-job images/sec, inputs resident in HBM (netcuda_forward_device, CUDA events on the launching stream, barrier +
               synchronize on both sides, max over ranks);
  e2e          the same metric through the reference-facing call, cuda::net_cuda::launch_forward(const std::vector<float>&) behind a
               net::net_abstract* (include/netAbstract.h:13): pageable host vector in, vector out, H2D and D2H inside the timed
               region; the pinned / asynchronous C-ABI calls are reported beside it;
  roofline     the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of the GEMM launches of one step / their summed CUDA-event
               durations (netcuda_profile_*, extra steps after the timed region), against the measured bf16 peak;
  cpu_baseline the CPU oracle (oracle/: the restatement the parity tests check against) on this box's host cores, bounded sample;
  configs      the other BASELINE.json configs on this GPU: C1 (MLP 784-128-64-10, batch 64), C2 (ViT-Tiny, bf16 and tf32),
               C4 (ViT-L/16-384), C5 (8 x 4096 INT8, batch sweep), each with its own roofline and cpu_baseline;
  N > 1        sharded_equals_single (gathered logits == the same global batch forwarded on one GPU, bit for bit), per-rank step
               times, the all-gather's own duration, and the strong-scaling split (global batch 1024).

`--impl reference` times the CPU implementation alone on the same config / metric (rank 0 only); it never loads libnetcuda.so.
Only the CPU legs (cpu_baseline, --impl reference) touch oracle/; the GPU path has no fallback.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))

WORKLOADS = {
    # ViT: (kind, preset, images per GPU per step, internal pass size, precision)
    "vit_base_16_224_b1024": ("vit", "vit_base_16_224", 1024, 512, "bf16"),
    "vit_tiny_16_224_b256": ("vit", "vit_tiny_16_224", 256, 256, "bf16"),
    "vit_tiny_16_224_b256_tf32": ("vit", "vit_tiny_16_224", 256, 256, "tf32"),
    "vit_large_16_384_b64": ("vit", "vit_large_16_384", 64, 32, "bf16"),
    # MLP: (kind, config name, samples per GPU per step, internal pass size, precision)
    "mlp_784_128_64_10_b64": ("mlp", "C1", 64, 64, "tf32"),
    "mlp_8x4096_int8_b16384": ("mlp", "C5", 16384, 16384, "int8"),
}
DEFAULT_WORKLOAD = "vit_base_16_224_b1024"
GEMM_LABELS = ("patch_embed", "qkv", "proj", "fc1", "fc2", "head", "mlp_layer", "mlp_layer_splitk")
STRONG_GLOBAL_BATCH = 1024


def host_threads() -> int:
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1, so omp_get_max_threads() is useless under it: the count is
    taken from the affinity mask and passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    # /opt/skills/guides/B200_PROFILING.md fallback figures
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic_per_launch(prof, pass_images):
    """dram__bytes_read + dram__bytes_write per GEMM launch from the committed ncu --set full capture (profiles/ncu_traffic.json),
    averaged over the GEMM launches of a step with the same weighting as roofline.achieved; None when no capture is committed or the
    capture was taken at another pass size."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    cap = json.load(open(path))
    if cap.get("pass_images") != pass_images:
        return None
    per = cap.get("per_launch_bytes", {})
    num = den = 0.0
    for k in ("qkv", "proj", "fc1", "fc2"):
        if k in per and k in prof:
            num += per[k] * prof[k]["launches"]
            den += prof[k]["launches"]
    return num / den if den else None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0: float, t1: float) -> dict:
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()


# =====================================================================================================================
# CPU legs -- the ONLY code in this file that touches oracle/ (cpu_baseline objects and --impl reference).  They import
# vit_presets (pure Python), never netcuda: the reference arm must not map the product library.
# =====================================================================================================================

def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import Oracle

    return Oracle()


def cpu_vit_rate(preset: str, budget_s: float, steps: int, warmup: int, threads: int, probe: bool = True):
    """The CPU oracle's ViT forward (fp32, OpenMP over images) on `threads` cores.  Each step forwards n images; n is sized from
    a one-image probe so that (steps + warmup) steps fit the budget (probe=False: one image per core, for nets whose single image
    already takes most of the budget -- ViT-L/16-384 is 382 GFLOP, about half a minute per image and core)."""
    import numpy as np
    import vit_presets as vp

    o = _oracle()
    cfg = vp.VIT_PRESETS[preset]
    flat = vp.vit_random_params(cfg, seed=0)
    rng = np.random.default_rng(1234)
    t_img = None
    n = threads
    if probe:
        x1 = rng.uniform(-1, 1, (1, 3 * cfg["image_size"] ** 2)).astype(np.float32)
        t = time.perf_counter()
        o.vit_forward(cfg, flat, x1, threads=1)
        t_img = time.perf_counter() - t  # one image on one core
        per_step = budget_s / max(steps + warmup, 1)
        # OpenMP runs over images, so a step costs at least t_img; use every core, several images per core if the budget allows
        n = threads * max(1, min(4, int(per_step / t_img)))
    x = rng.uniform(-1, 1, (n, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    for _ in range(warmup):
        o.vit_forward(cfg, flat, x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.vit_forward(cfg, flat, x, threads=threads)
    dt = time.perf_counter() - t0
    return dict(value=n * steps / dt, images_per_step=n, cores=threads, seconds=dt, ms_per_step=dt / steps * 1e3,
                one_image_one_core_s=t_img)


def cpu_vit_baseline(preset: str, budget_s: float, threads: int, probe: bool = True) -> dict:
    r = cpu_vit_rate(preset, budget_s, steps=1, warmup=0, threads=threads, probe=probe)
    return {"value": r["value"], "unit": "images/sec", "cores": r["cores"], "kind": "port",
            "sample": f"{r['images_per_step']} images of the workload, one pass, fp32 oracle (oracle/oracle_vit.c), "
                      f"{r['seconds']:.1f} s on {r['cores']} of {os.cpu_count()} host cores"}


def cpu_mlp_baselines(name: str, threads: int, budget_s: float = 8.0) -> dict:
    """MLP configs: (a) the reference's own unchanged src/netFPGA.cpp over the OpenCL shim (oracle/_ref), one sample per
    launch_forward call and single-threaded exactly like the reference (src/netFPGA.cpp:239-290) -- kind "reference-shim";
    (b) the oracle port, batched, on all cores (for C5: the INT8 Q1.7 port the bit-exact tests check against)."""
    import numpy as np
    import vit_presets as vp

    o = _oracle()
    from oracle import Reference  # (sys.path was extended by _oracle())

    cfg = vp.MLP_C1 if name == "C1" else vp.MLP_C5
    npl, n_ins = cfg["npl"], cfg["n_ins"]
    rng = np.random.default_rng(1234)
    out = {}
    w, b = vp.mlp_reference_rule_params(npl, n_ins, seed=1)
    if name == "C5":
        w = (w * np.float32(1.0 / np.sqrt(n_ins))).astype(np.float32)
    if Reference.available():
        ref = Reference()
        h = ref.create(npl, n_ins, w, b)
        x = rng.uniform(-1, 1, (1, n_ins)).astype(np.float32)
        ref.forward(h, x, n_ins, npl[-1])
        t0 = time.perf_counter()
        calls = 0
        while True:
            ref.forward(h, x, n_ins, npl[-1])  # one sample per call: the reference's launch_forward contract
            calls += 1
            dt = time.perf_counter() - t0
            if dt > budget_s / 2 or calls >= 20000:
                break
        ref.destroy(h)
        out["reference_shim"] = {"value": calls / dt, "unit": "samples/sec", "cores": 1, "kind": "reference-shim",
                                 "sample": f"{calls} launch_forward calls of one sample each through the unchanged src/netFPGA.cpp over the "
                                           f"OpenCL shim (fp32, single-threaded like the reference), {dt:.1f} s"}
    if name == "C1":
        n = 64
        x = rng.uniform(-1, 1, (n, n_ins)).astype(np.float32)
        o.mlp_forward(x, w, b, npl, n_ins, threads=threads)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 1.0:
            o.mlp_forward(x, w, b, npl, n_ins, threads=threads)
            reps += 1
        dt = time.perf_counter() - t0
        out["port"] = {"value": n * reps / dt, "unit": "samples/sec", "cores": threads, "kind": "port",
                       "sample": f"{reps} batches of 64 samples, fp32 oracle (oracle/oracle_mlp.c), {dt:.1f} s"}
    else:
        wq, bq = vp.mlp_int8_params(npl, n_ins)
        n = 2 * threads
        xq = rng.integers(-128, 128, (n, n_ins), dtype=np.int8)
        o.mlp_forward_i8(xq[:threads], wq, bq, npl, n_ins, threads=threads)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < budget_s / 2:
            o.mlp_forward_i8(xq, wq, bq, npl, n_ins, threads=threads)
            reps += 1
        dt = time.perf_counter() - t0
        out["port"] = {"value": n * reps / dt, "unit": "samples/sec", "cores": threads, "kind": "port",
                       "sample": f"{reps} batches of {n} samples, INT8 Q1.7 oracle (oracle/oracle_mlp.c), {dt:.1f} s"}
    return out


def run_reference(args, kind, name, per_gpu):
    """`--impl reference`: the CPU implementation of the path alone, all host cores, bounded sample; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = host_threads()
    if kind == "vit":
        r = cpu_vit_rate(name, budget_s=120.0, steps=args.steps, warmup=args.warmup, threads=threads)
        value, ms, unit = r["value"], r["ms_per_step"], "images/sec"
        sample = (f"{r['images_per_step']} images of the workload per step (the GPU arm forwards {per_gpu} per GPU: same net, same rate "
                  f"metric, smaller step), fp32, OpenMP over images")
        base = {"value": value, "unit": unit, "cores": r["cores"], "kind": "port", "sample": sample}
        cfg = {"workload": args.workload, "net": name, "images_per_step": r["images_per_step"], "same_config_as_gpu_arm": False,
               "note": "the reference ships no ViT and no device kernel (SURVEY.md s.0); this is the CPU oracle port of the path"}
    else:
        b = cpu_mlp_baselines(name, threads, budget_s=30.0)
        pick = b.get("reference_shim") or b["port"]
        value, unit = pick["value"], "samples/sec"
        ms = 1e3 / value
        base = pick
        cfg = {"workload": args.workload, "net": name, "samples_per_step": 1 if pick["kind"] == "reference-shim" else per_gpu,
               "same_config_as_gpu_arm": False, "also": b}
    # the MLP configs through the unchanged reference host runtime (oracle/_ref), whatever the headline workload is
    configs = {}
    if kind == "vit":
        for c in ("C1", "C5"):
            try:
                configs[c] = cpu_mlp_baselines(c, threads, budget_s=8.0)
            except Exception as e:  # noqa: BLE001 (a missing _ref must not void the headline line)
                configs[c] = {"error": str(e)}
    line = {
        "impl": "reference", "metric": unit.replace("/sec", "") + "/sec", "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": threads, "configs": configs,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# GPU legs
# =====================================================================================================================

def time_steps(torch, stream, fn, steps, warmup):
    """CUDA-event time (ms) of `steps` calls of fn() on `stream`, after `warmup` untimed ones, synchronised on both sides."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def profile_per_kernel(net, fwd, steps):
    net.profile_enable(True)
    for _ in range(steps):
        fwd()
    prof = net.profile_read()
    net.profile_enable(False)
    return prof


def per_kernel_table(prof, steps):
    return {k: {"launches": v["launches"] // steps, "ms_per_step": round(v["ms"] / steps, 4),
                "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] > 0 and v["ms"] > 0 else None,
                "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}


def gemm_roofline(prof, peaks, extra=None):
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k in GEMM_LABELS)
    gemm_flops = sum(v["flops"] for k, v in prof.items() if k in GEMM_LABELS)
    gemm_launches = sum(v["launches"] for k, v in prof.items() if k in GEMM_LABELS)
    all_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak = peaks["sustained"]  # the GEMMs are timed inside a long step, under the power cap
    r = {"bound": "tensor", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
         "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
         "frac_of_burst_peak": round(achieved / peaks["burst"], 4),
         "flops_per_launch": gemm_flops / max(gemm_launches, 1), "ms_per_launch": gemm_ms / max(gemm_launches, 1),
         "algorithmic_bytes_per_launch": sum(v["bytes"] for k, v in prof.items() if k in GEMM_LABELS) / max(gemm_launches, 1),
         "share_of_step": round(gemm_ms / all_ms, 4) if all_ms > 0 else None, "traffic": None}
    if extra:
        r.update(extra)
    return r


def bench_vit_config(nc, torch, dev, local, preset, per_gpu, pass_size, precision, steps, warmup, peaks, threads, cpu_budget_s):
    """One ViT config, device-resident, on this GPU: value, whole-step roofline, per-kernel table, cpu_baseline."""
    import vit_presets as vp

    cfg = vp.VIT_PRESETS[preset]
    net = nc.Net.vit(cfg, device=local, max_batch=pass_size, precision=nc.PRECISIONS[precision])
    net.upload_vit(vp.vit_random_params(cfg, seed=0))
    gen = torch.Generator(device=dev).manual_seed(1234)
    x = torch.rand((per_gpu, net.n_in), generator=gen, device=dev) * 2 - 1
    y = torch.empty((per_gpu, net.n_out), device=dev)
    stream = torch.cuda.current_stream()
    fwd = lambda: net.forward_device(x, y, per_gpu, stream)  # noqa: E731
    ms = time_steps(torch, stream, fwd, steps, warmup)
    value = per_gpu * steps / (ms * 1e-3)
    prof_steps = min(steps, 3)
    prof = profile_per_kernel(net, fwd, prof_steps)
    torch.cuda.synchronize()
    flops = net.flops_per_sample
    net.close()
    del x, y
    tf = value * flops / 1e12
    # tf32 MMAs run at half the bf16 rate: its fraction is still quoted against the measured bf16 peak, and says so
    roof = gemm_roofline(prof, peaks, {"kernel": f"gemm_tn_tcgen05_kernel<{precision}> (all GEMM launches of a step)",
                                       "whole_step_tflops": round(tf, 2), "whole_step_frac_of_burst_peak": round(tf / peaks["burst"], 4),
                                       "per_kernel": per_kernel_table(prof, prof_steps)})
    out = {"workload": f"{preset}, {per_gpu} images per step, pass size {pass_size}, {precision}", "metric": "images/sec", "value": value,
           "unit": "images/sec", "ms_per_step": ms / steps, "steps": steps, "dtype": precision, "flops_per_image": flops, "roofline": roof}
    if cpu_budget_s > 0:
        out["cpu_baseline"] = cpu_vit_baseline(preset, cpu_budget_s, threads, probe=preset != "vit_large_16_384")
    return out


def bench_c1(nc, torch, dev, local, threads, with_cpu):
    """Config C1: MLP 784-128-64-10, batch 64 -- the plumbing config (launch / PCIe latency bound: no roofline claim).
    Through the C++ class behind net::net_abstract* (one sample per call = the reference's contract, and the batch of 64), and
    device-resident through the C ABI, in the bit-exact fp32 precision (the class default) and on tensor cores."""
    import numpy as np
    import vit_presets as vp

    npl, n_ins = vp.MLP_C1["npl"], vp.MLP_C1["n_ins"]
    w, b = vp.mlp_reference_rule_params(npl, n_ins, seed=1)
    x = np.random.default_rng(1234).uniform(-1, 1, (64, n_ins)).astype(np.float32)
    out = {"workload": "MLP 784-128-64-10, batch 64", "metric": "samples/sec", "unit": "samples/sec", "precisions": {}}
    stream = torch.cuda.current_stream()
    for prec in ("fp32", "tf32", "bf16"):
        h = nc.HostNet.mlp(npl, n_ins, w, b, precision=nc.PRECISIONS[prec], device=local)
        s1, _ = h.time_launch_forward(x[:1], reps=200)
        s64, _ = h.time_launch_forward(x, reps=200)
        h.close()
        net = nc.Net.mlp(npl, n_ins, precision=nc.PRECISIONS[prec], device=local)
        net.upload_mlp(w, b)
        dx = torch.from_numpy(x).to(dev)
        dy = torch.empty((64, npl[-1]), device=dev)
        ms = time_steps(torch, stream, lambda: net.forward_device(dx, dy, 64, stream), 200, 20)
        net.close()
        out["precisions"][prec] = {"launch_forward_one_sample_us": round(s1 * 1e6, 1), "launch_forward_batch64_us": round(s64 * 1e6, 1),
                                   "device_resident_batch64_us": round(ms / 200 * 1e3, 1),
                                   "samples_per_sec_launch_forward_batch64": round(64 / s64, 0),
                                   "samples_per_sec_device_resident": round(64 / (ms / 200 * 1e-3), 0)}
    p = out["precisions"]["fp32"]
    out["value"] = p["samples_per_sec_launch_forward_batch64"]
    out["dtype"] = "f32"
    out["api"] = "net_cuda::launch_forward(std::vector) through net::net_abstract*, 64 samples per call, fp32 (bit-equal to the oracle)"
    out["roofline"] = {"bound": "latency", "note": "218 kFLOP per sample: launch- and PCIe-latency bound, no roofline claim (SURVEY.md s.8d)"}
    if with_cpu:
        out["cpu_baseline"] = cpu_mlp_baselines("C1", threads)
    return out


def measure_int8_peak(torch, dev):
    """Dense INT8 tensor-core rate of this GPU, measured in this run with the vendor library (torch._int_mm -> cuBLASLt) on the
    layer shape of config C5 at its largest batch; TOP/s, best of 10.  None if the library call is unavailable."""
    try:
        a = torch.randint(-128, 128, (16384, 4096), device=dev, dtype=torch.int32).to(torch.int8)
        b = torch.randint(-128, 128, (4096, 4096), device=dev, dtype=torch.int32).to(torch.int8).t()  # column-major B, as cuBLASLt wants
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 16384 * 4096 * 4096 / (best * 1e-3) / 1e12
    except Exception:  # noqa: BLE001
        return None


def bench_c5(nc, torch, dev, local, peaks, threads, with_cpu):
    """Config C5: 8 x 4096 INT8 (Q1.7), bit-exact, batch sweep.  Below ~128 samples a forward is weight streaming: 134,217,728 bytes of
    int8 weights against a few KB of activations -> GB/s of weights vs the measured HBM peak.  Above, TOP/s (2 x 8 x 4096^2 int-ops
    per sample) vs the INT8 peak measured in this run.  Weights (128 MiB) exceed the 126 MB L2: no flush between iterations."""
    import numpy as np
    import vit_presets as vp

    npl, n_ins = vp.MLP_C5["npl"], vp.MLP_C5["n_ins"]
    wq, bq = vp.mlp_int8_params(npl, n_ins)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, device=local, max_batch=16384)
    net.upload_mlp_i8(wq, bq)
    weight_bytes = float(sum(a * b for a, b in zip([n_ins] + npl[:-1], npl)))
    ops_per_sample = 2.0 * weight_bytes
    int8_peak = measure_int8_peak(torch, dev)
    stream = torch.cuda.current_stream()
    sweep = {}
    for batch in (1, 4, 8, 16, 32, 64, 128, 1024, 16384):  # (kernels: mma.sync stream <= 16, tcgen05 split-K cluster stream 17..128, tcgen05 GEMMs above)
        xq = torch.randint(-128, 128, (batch, n_ins), device=dev, dtype=torch.int32).to(torch.int8)
        yq = torch.empty((batch, npl[-1]), dtype=torch.int32, device=dev)
        steps = 200 if batch <= 1024 else 20
        ms = time_steps(torch, stream, lambda: net.forward_device_i8(xq, yq, batch, stream), steps, 10) / steps
        gbs = (weight_bytes + 2.0 * batch * n_ins) / (ms * 1e-3) / 1e9
        tops = ops_per_sample * batch / (ms * 1e-3) / 1e12
        row = {"us_per_forward": round(ms * 1e3, 1), "samples_per_sec": round(batch / (ms * 1e-3), 0)}
        if batch <= 128:
            row["roofline"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(gbs / peaks["hbm"], 4)}
        else:
            row["roofline"] = {"bound": "tensor", "achieved": round(tops, 3), "peak": round(int8_peak, 3) if int8_peak else None, "unit": "TOP/s",
                               "frac": round(tops / int8_peak, 4) if int8_peak else None,
                               "peak_source": "torch._int_mm (cuBLASLt) 16384x4096x4096, best of 10, measured in this run"}
        sweep[str(batch)] = row
    # e2e of the largest batch through the host-buffer call (int8 in, int32 out)
    xq = np.random.default_rng(5).integers(-128, 128, (16384, n_ins), dtype=np.int8)
    yq_host = np.empty((16384, npl[-1]), dtype=np.int32)
    for _ in range(5):  # (first touch of the output pages, and the first use of each of the library's four in-flight output slots, stay outside the clock)
        net.forward_i8(xq, out=yq_host)
    t0 = time.perf_counter()
    for _ in range(3):
        net.forward_i8(xq, out=yq_host)
    e2e = 3 * 16384 / (time.perf_counter() - t0)
    net.close()
    out = {"workload": "MLP 8 x 4096, INT8 Q1.7, bit-exact, batch sweep", "metric": "samples/sec", "unit": "samples/sec", "dtype": "s8",
           "value": sweep["16384"]["samples_per_sec"], "algorithmic_bytes_per_forward": weight_bytes, "int_ops_per_sample": ops_per_sample,
           "int8_peak_tops_measured": int8_peak, "sweep": sweep, "roofline": sweep["16384"]["roofline"],
           "e2e": {"value": e2e, "unit": "samples/sec", "api": "netcuda_forward_i8, pageable host buffers, batch 16384"}}
    if with_cpu:
        out["cpu_baseline"] = cpu_mlp_baselines("C5", threads)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="netcuda", choices=["netcuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU (default); strong: a global batch of 1024 split over the GPUs")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (default: the workload's)")
    ap.add_argument("--max-batch", type=int, default=0, help="samples per internal pass")
    ap.add_argument("--gemm-variant", type=int, default=0, help="netcuda_set_gemm_variant (A/B measurements; 0 = product path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (C1, C2, C4, C5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "netcuda" else args.warmup
    kind, name, per_gpu, max_batch, precision = WORKLOADS[args.workload]
    per_gpu = args.batch or per_gpu
    max_batch = args.max_batch or max_batch
    if args.impl == "reference":
        run_reference(args, kind, name, per_gpu)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))

    import numpy as np
    import torch
    import torch.distributed as dist
    import netcuda as nc  # ImportError if the native library is missing: there is no fallback
    import vit_presets as vp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl netcuda needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    threads = host_threads()
    peaks = load_peaks()
    if args.scaling == "strong":
        per_gpu = max(1, -(-STRONG_GLOBAL_BATCH // world))
        max_batch = min(max_batch, per_gpu)

    # ---- the net and this rank's shard of the global batch ----------------------------------------------
    int8 = precision == "int8"
    if kind == "vit":
        cfg = vp.VIT_PRESETS[name]
        flat = vp.vit_random_params(cfg, seed=0)
        net = nc.Net.vit(cfg, device=local, max_batch=max_batch, precision=nc.PRECISIONS[precision])
        net.upload_vit(flat)
        unit = "images/sec"
    else:
        mcfg = vp.MLP_C1 if name == "C1" else vp.MLP_C5
        net = nc.Net.mlp(mcfg["npl"], mcfg["n_ins"], precision=nc.PRECISIONS[precision], device=local, max_batch=max_batch)
        if int8:
            net.upload_mlp_i8(*vp.mlp_int8_params(mcfg["npl"], mcfg["n_ins"]))
        else:
            net.upload_mlp(*vp.mlp_reference_rule_params(mcfg["npl"], mcfg["n_ins"], seed=1))
        unit = "samples/sec"
    if args.gemm_variant:
        net.set_gemm_variant(args.gemm_variant)
    n_in, n_out = net.n_in, net.n_out

    def make_shard(r, n):
        """Shard r of the global batch: uniform [-1, 1) (MIN/MAX_RANGE of def/defines.h:11-12), the same values whichever rank draws it."""
        gen = torch.Generator(device=dev).manual_seed(1234 + r)
        if int8:
            return torch.randint(-128, 128, (n, n_in), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)
        return torch.rand((n, n_in), generator=gen, device=dev) * 2 - 1

    out_dtype = torch.int32 if int8 else torch.float32
    fwd_dev = net.forward_device_i8 if int8 else net.forward_device
    x = make_shard(rank, per_gpu)
    y = [torch.empty((per_gpu, n_out), device=dev, dtype=out_dtype) for _ in range(2)]
    gathered = [torch.empty((world * per_gpu, n_out), device=dev, dtype=out_dtype) for _ in range(2)] if world > 1 else None
    from netcuda.sharding import shard_bounds

    lo, hi = shard_bounds(world * per_gpu, world, rank)  # this rank's slice of the global batch
    assert hi - lo == per_gpu
    # kernels and timing events run on one explicit side stream, the logits all-gather on a second one (it overlaps the next step)
    torch.cuda.synchronize()  # inputs were generated on the default stream
    stream, cstream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    fwd_done = [torch.cuda.Event(), torch.cuda.Event()]
    gather_done = [torch.cuda.Event(), torch.cuda.Event()]
    step_no = [0]

    def step():
        b = step_no[0] & 1
        step_no[0] += 1
        if world > 1:
            stream.wait_event(gather_done[b])  # y[b] is free again: the gather of two steps ago has read it
        fwd_dev(x, y[b], per_gpu, stream)
        if world > 1:
            fwd_done[b].record(stream)
            cstream.wait_event(fwd_done[b])
            with torch.cuda.stream(cstream):
                dist.all_gather_into_tensor(gathered[b], y[b])  # NCCL all-gather of the logits
                gather_done[b].record(cstream)

    def drain():
        if world > 1:
            stream.wait_event(gather_done[0])
            stream.wait_event(gather_done[1])

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- device-resident throughput ---------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    drain()
    fence()
    l0 = net.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    drain()  # the timed region ends when the last gathered logits have landed
    e1.record(stream)
    fence()
    t_b = time.perf_counter()
    launches = net.launches - l0
    wall_ms = (t_b - t_a) * 1e3
    ev_ms = e0.elapsed_time(e1)
    if abs(ev_ms - wall_ms) > 0.05 * wall_ms + 2.0 + (3.0 * world if world > 1 else 0.0):  # the host waits on the device here: the clocks must agree
        raise SystemExit(f"timing inconsistency: CUDA events {ev_ms:.2f} ms vs host clock {wall_ms:.2f} ms around the same region")
    ms = torch.tensor([ev_ms], device=dev, dtype=torch.float64)
    per_rank_ms = [ev_ms / args.steps]
    if world > 1:
        all_ms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(all_ms, ms)
        per_rank_ms = [float(t.item()) / args.steps for t in all_ms]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * per_gpu * args.steps / (ms_total * 1e-3)

    extra = {}
    if world > 1:
        # the collective alone (same buffers), and the forward alone, so that the 1 -> N loss can be attributed
        ag_ms = time_steps(torch, cstream, lambda: _gather_on(torch, dist, cstream, gathered[0], y[0]), 20, 3) / 20
        fwd_ms = time_steps(torch, stream, lambda: fwd_dev(x, y[0], per_gpu, stream), min(args.steps, 5), 1) / min(args.steps, 5)
        red = torch.tensor([ag_ms, fwd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        extra["per_rank_ms_per_step"] = [round(v, 3) for v in per_rank_ms]
        extra["all_gather_ms"] = round(float(red[0].item()), 4)
        extra["all_gather_bytes_out_per_rank"] = world * per_gpu * n_out * 4
        extra["forward_only_ms_per_step_max_rank"] = round(float(red[1].item()), 3)
        # ---- bit-identity: the gathered logits of one step == the same global batch forwarded on ONE GPU (rank 0) ----
        step()
        drain()
        fence()
        flag = torch.ones(1, device=dev, dtype=torch.int32)
        if rank == 0:
            xg = torch.cat([x] + [make_shard(r, per_gpu) for r in range(1, world)])
            yg = torch.empty((world * per_gpu, n_out), device=dev, dtype=out_dtype)
            fwd_dev(xg, yg, world * per_gpu, stream)
            torch.cuda.synchronize()
            flag[0] = int(torch.equal(yg, gathered[(step_no[0] - 1) & 1]))
            extra["sharded_max_abs_diff_vs_single"] = float((yg.double() - gathered[(step_no[0] - 1) & 1].double()).abs().max().item())
            del xg, yg
        dist.broadcast(flag, 0)
        extra["sharded_equals_single"] = bool(flag.item())
        # ---- strong scaling: a global batch of 1024 split over the GPUs (ceil(B / G) samples per GPU) ----
        if args.scaling == "weak" and kind == "vit":
            sp = -(-STRONG_GLOBAL_BATCH // world)
            xs, ys = x[:sp], y[0][:sp]
            gs = gathered[0][: world * sp]

            def strong_step():
                fwd_dev(xs, ys, sp, stream)
                dist.all_gather_into_tensor(gs, ys)

            s_ms = torch.tensor([time_steps(torch, stream, strong_step, max(args.steps, 10), 3) / max(args.steps, 10)], device=dev, dtype=torch.float64)
            dist.all_reduce(s_ms, op=dist.ReduceOp.MAX)
            extra["strong_scaling"] = {"global_batch": world * sp, "images_per_gpu": sp, "ms_per_step": round(float(s_ms.item()), 4),
                                       "value": world * sp / (float(s_ms.item()) * 1e-3), "unit": unit,
                                       "note": "same net and kernels; one pass per step, logits all-gather inside the step"}
        fence()

    # ---- per-kernel CUDA-event times (extra steps, events around every launch) ----------------------------
    prof_steps = min(args.steps, 3)
    prof = profile_per_kernel(net, lambda: fwd_dev(x, y[0], per_gpu, stream), prof_steps)
    torch.cuda.synchronize()

    # ---- end to end -------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e and not int8:
        hx = torch.empty((per_gpu, n_in), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hy = torch.empty((per_gpu, n_out), dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 10))
        # (1) the reference-facing call: net::net_abstract::launch_forward(const std::vector<float>&) on a cuda::net_cuda object
        #     (its own netcuda handle, same weights): pageable vector in, vector out by value, H2D + kernels + D2H inside the clock
        def class_call():
            if kind == "vit":
                host = nc.HostNet.vit(cfg, flat, device=local, max_batch=max_batch, precision=nc.PRECISIONS[precision])
            else:
                host = nc.HostNet.mlp(mcfg["npl"], mcfg["n_ins"], *vp.mlp_reference_rule_params(mcfg["npl"], mcfg["n_ins"], seed=1),
                                      precision=nc.PRECISIONS[precision], device=local, max_batch=max_batch)
            fence()
            s_call, y_call = host.time_launch_forward(hx.numpy(), reps=e2e_steps)
            host.close()
            return torch.tensor([s_call], device=dev, dtype=torch.float64), y_call

        dt_class, y_class = class_call()
        # (1b) the same call on a net built with net_cuda_options::pin_inputs (NETCUDA_PIN_INPUTS=1): the caller's vector is page-locked
        #      by the first (warm-up) call and every later call is DMA'd straight from it -- no staging copy
        os.environ["NETCUDA_PIN_INPUTS"] = "1"
        try:
            dt_class_pin, y_class_pin = class_call()
        finally:
            del os.environ["NETCUDA_PIN_INPUTS"]
        # (2) the C ABI with page-locked buffers: blocking netcuda_forward, and netcuda_submit / netcuda_wait with two calls in flight
        for _ in range(2):
            net.forward_into(hx, hy)
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            net.forward_into(hx, hy)  # synchronous: returns when the logits are in host memory
        torch.cuda.synchronize()
        dt_sync = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        hx2 = torch.empty_like(hx, pin_memory=True)
        hx2.copy_(hx)
        hy2 = torch.empty_like(hy, pin_memory=True)
        bufs = ((hx, hy), (hx2, hy2))
        net.wait(net.submit(hx, hy))
        fence()
        t0 = time.perf_counter()
        prev = None
        for i in range(e2e_steps):
            tk = net.submit(*bufs[i & 1])
            if prev is not None:
                net.wait(prev)
            prev = tk
        net.wait(prev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            for t in (dt, dt_sync, dt_class, dt_class_pin):
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = world * per_gpu
        e2e = {"value": total / float(dt_class.item()), "unit": unit,
               "h2d_bytes_per_step": total * n_in * 4, "d2h_bytes_per_step": total * n_out * 4, "steps": e2e_steps,
               "api": "net::net_abstract::launch_forward(const std::vector<float>&) on cuda::net_cuda (pageable std::vector in, std::vector out; "
                      "staging copy, H2D, kernels and D2H inside the timed region)",
               "pin_inputs_value": total / float(dt_class_pin.item()),
               "pin_inputs_api": "the same launch_forward(std::vector) on a net with net_cuda_options::pin_inputs (NETCUDA_PIN_INPUTS=1): the "
                                 "vector is page-locked on first sight, later calls DMA from it in place",
               "pinned_async_value": total * e2e_steps / float(dt.item()),
               "pinned_async_api": "netcuda_submit / netcuda_wait, two calls in flight, page-locked buffers (H2D of call i+1 overlaps the kernels of call i)",
               "pinned_blocking_value": total * e2e_steps / float(dt_sync.item()),
               "pinned_blocking_api": "netcuda_forward, page-locked buffers (2-slot staged H2D inside the call)"}
        assert bool((hy2 == hy).all())
        e2e["max_abs_diff_vs_device_path"] = float((hy.to(dev) - y[0]).abs().max().item())  # same kernels, same inputs: must be 0
        e2e["class_max_abs_diff_vs_device_path"] = float(max(np.abs(y_class - y[0].cpu().numpy()).max(), np.abs(y_class_pin - y[0].cpu().numpy()).max()))
        del hx, hy, hx2, hy2
    t_c = time.perf_counter()

    # ---- BASELINE config 4 is the 8-GPU config: measured at every N (64 images per GPU); C1 / C2 / C5 on one GPU ----
    flops_per_sample = net.flops_per_sample
    net.close()
    del x, y
    torch.cuda.empty_cache()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    configs = {}
    if not args.no_configs and args.workload == DEFAULT_WORKLOAD and args.scaling == "weak":
        with_cpu = rank == 0 and not args.no_cpu_baseline and world == 1
        c4 = bench_vit_config(nc, torch, dev, local, "vit_large_16_384", 64, 32, "bf16", 5, 3, peaks, threads, 25.0 if with_cpu else 0.0)
        if world > 1:
            v = torch.tensor([c4["ms_per_step"]], device=dev, dtype=torch.float64)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            c4["ms_per_step"] = float(v.item())
            c4["value"] = world * 64 / (c4["ms_per_step"] * 1e-3)
            c4["n_gpus"] = world
            c4["note"] = "64 images per GPU (= BASELINE config 4's 512 on 8 GPUs); value = whole job, max over ranks; no gather in this leg"
        configs["C4"] = c4
        if world == 1:
            configs["C2"] = bench_vit_config(nc, torch, dev, local, "vit_tiny_16_224", 256, 256, "bf16", 20, 5, peaks, threads, 6.0 if with_cpu else 0.0)
            configs["C2_tf32"] = bench_vit_config(nc, torch, dev, local, "vit_tiny_16_224", 256, 256, "tf32", 20, 5, peaks, threads, 0.0)
            configs["C1"] = bench_c1(nc, torch, dev, local, threads, with_cpu)
            configs["C5"] = bench_c5(nc, torch, dev, local, peaks, threads, with_cpu)
    t_d = time.perf_counter()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    clocks = sampler.window(t_a, t_b)
    sampler.stop()
    tf = value / world * flops_per_sample / 1e12
    roofline = gemm_roofline(prof, peaks, {
        "kernel": f"gemm_tn_tcgen05_kernel<{precision}> (all GEMM launches of a step)",
        "traffic": ncu_traffic_per_launch(prof, min(max_batch, per_gpu)) if kind == "vit" else None,
        "whole_step_tflops": round(tf, 2), "whole_step_frac_of_burst_peak": round(tf / peaks["burst"], 4),
        "whole_step_frac_of_sustained_peak": round(tf / peaks["sustained"], 4),
        "per_kernel": per_kernel_table(prof, prof_steps)})

    cpu_baseline = None
    if not args.no_cpu_baseline:
        if kind == "vit":
            cpu_baseline = cpu_vit_baseline(name, 20.0, threads)
        else:
            b = cpu_mlp_baselines(name, threads)
            cpu_baseline = b.get("reference_shim") or b["port"]

    line = {
        "metric": unit, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": precision if precision != "int8" else "s8",
        "data": "synthetic",
        "config": {"workload": args.workload, "net": name, "samples_per_gpu_per_step": per_gpu, "global_batch": world * per_gpu,
                   "pass_size": max_batch, "weights": "random-init, replicated per GPU",
                   "sharding": f"batch x{world}, logits all-gather on a side stream (overlaps the next step)" if world > 1 else "single GPU",
                   "l2": f"inputs ({per_gpu * n_in * (1 if int8 else 4) >> 20} MiB per GPU) and per-pass activations exceed the 126 MB L2; no flush needed",
                   "flops_per_sample": flops_per_sample},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "host_cores": threads,
    }
    line.update(extra)
    if e2e is not None:
        line["e2e"] = e2e
    if configs:
        line["configs"] = configs
    line["wall_s"] = {"timed_region": round(t_b - t_a, 3), "profile_and_e2e": round(t_c - t_b, 3), "other_configs": round(t_d - t_c, 3)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _gather_on(torch, dist, cstream, out, src):
    with torch.cuda.stream(cstream):
        dist.all_gather_into_tensor(out, src)


if __name__ == "__main__":
    main()
