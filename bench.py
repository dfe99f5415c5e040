#!/usr/bin/env python
"""bench.py -- images/sec of the batched forward pass through the netCUDA backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl netcuda|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of one batch of synthetic images (ViT-B/16-224, 1024 images per GPU,
random-init weights) through the net.  One process per GPU; every rank forwards its own shard of the
global batch with replicated weights (no data-path collective), then the logits are gathered with one
NCCL all-gather (N > 1), as BASELINE.json's north_star describes.

One JSON line on rank 0:
  value      whole-job images/sec, inputs resident in HBM (netcuda_forward_device, CUDA events on the
             launching stream, barrier + synchronize on both sides, max over ranks);
  e2e        the same metric through the host-buffer call the C++ class makes (netcuda_forward): pinned
             host input -> H2D -> kernels -> D2H of the logits, all inside the timed region;
  roofline   the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of the GEMM launches of one step /
             their summed CUDA-event durations (netcuda_profile_*, measured in extra steps after the
             timed region), against the measured bf16 peak of MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle (oracle/, the restatement the parity tests check against) timed on this
             box's host cores on a bounded sample of the same workload.

`--impl reference` times the CPU implementation alone on the same config/metric (rank 0 only).
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
    # name: (preset, images per GPU per step, internal pass size)
    "vit_base_16_224_b1024": ("vit_base_16_224", 1024, 512),
    "vit_tiny_16_224_b256": ("vit_tiny_16_224", 256, 256),
    "vit_large_16_384_b64": ("vit_large_16_384", 64, 32),
}
DEFAULT_WORKLOAD = "vit_base_16_224_b1024"
GEMM_LABELS = ("patch_embed", "qkv", "proj", "fc1", "fc2", "head")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    # /opt/skills/guides/B200_PROFILING.md fallback figures
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic_per_launch(prof, prof_steps, pass_images):
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


def cpu_forward_rate(preset: str, budget_s: float, steps: int, warmup: int):
    """Times the CPU oracle (the restatement of the path the parity tests use as checker) on this box's
    cores.  Each step forwards `n` images (OpenMP over images); n is sized from a one-image probe so
    that (steps + warmup) steps fit the budget."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import netcuda as nc
    from oracle import Oracle

    o = Oracle()
    cfg = nc.VIT_PRESETS[preset]
    flat = nc.vit_random_params(cfg, seed=0)
    rng = np.random.default_rng(1234)
    threads = o.threads
    x1 = rng.uniform(-1, 1, (1, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    t = time.perf_counter()
    o.vit_forward(cfg, flat, x1, threads=1)
    t_img = time.perf_counter() - t  # one image on one core
    per_step = budget_s / max(steps + warmup, 1)
    # OpenMP runs over images, so a step costs at least t_img; use every core, several images per core if the budget allows
    n = threads * max(1, min(4, int(per_step / t_img)))
    used = threads
    x = rng.uniform(-1, 1, (n, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    for _ in range(warmup):
        o.vit_forward(cfg, flat, x, threads=used)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.vit_forward(cfg, flat, x, threads=used)
    dt = time.perf_counter() - t0
    return dict(value=n * steps / dt, images_per_step=n, cores=used, seconds=dt, ms_per_step=dt / steps * 1e3)


def run_reference(args, preset, per_gpu):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.workload
    r = cpu_forward_rate(preset, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    sample = f"{r['images_per_step']} images of the workload per step, fp32, OpenMP over images"
    line = {
        "impl": "reference", "metric": "images/sec", "value": r["value"], "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg_name, "images_per_step": r["images_per_step"],
                   "note": "the reference ships no ViT and no device kernel (SURVEY.md s.0); this is the CPU oracle port of the path"},
        "cpu_baseline": {"value": r["value"], "unit": "images/sec", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="netcuda", choices=["netcuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the workload's)")
    ap.add_argument("--max-batch", type=int, default=0, help="images per internal pass")
    ap.add_argument("--gemm-variant", type=int, default=0, help="netcuda_set_gemm_variant (A/B measurements; 0 = product path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "netcuda" else args.warmup
    preset, per_gpu, max_batch = WORKLOADS[args.workload]
    per_gpu = args.batch or per_gpu
    max_batch = args.max_batch or max_batch
    if args.impl == "reference":
        run_reference(args, preset, per_gpu)
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
    cfg = nc.VIT_PRESETS[preset]

    net = nc.Net.vit(cfg, device=local, max_batch=max_batch)
    net.upload_vit(nc.vit_random_params(cfg, seed=0))
    if args.gemm_variant:
        net.set_gemm_variant(args.gemm_variant)
    n_in, n_out = net.n_in, net.n_out
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((per_gpu, n_in), generator=gen, device=dev) * 2 - 1  # uniform [-1, 1): MIN/MAX_RANGE of def/defines.h:11-12
    y = torch.empty((per_gpu, n_out), device=dev)
    from netcuda.sharding import gather_outputs, shard_bounds

    lo, hi = shard_bounds(world * per_gpu, world, rank)  # this rank's slice of the global batch
    assert hi - lo == per_gpu
    # everything (kernels, the logits all-gather, the timing events) runs on one explicit side stream
    torch.cuda.synchronize()  # inputs were generated on the default stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    gathered = [None]

    def step():
        net.forward_device(x, y, per_gpu, stream)
        gathered[0] = gather_outputs(y, world * per_gpu, world)  # NCCL all-gather of the logits (no-op on one GPU)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- device-resident throughput ---------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    fence()
    l0 = net.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    fence()
    t_b = time.perf_counter()
    launches = net.launches - l0
    wall_ms = (t_b - t_a) * 1e3
    ev_ms = e0.elapsed_time(e1)
    if abs(ev_ms - wall_ms) > 0.05 * wall_ms + 2.0:  # the host waits on the device here, so the two clocks must agree
        raise SystemExit(f"timing inconsistency: CUDA events {ev_ms:.2f} ms vs host clock {wall_ms:.2f} ms around the same region")
    ms = torch.tensor([ev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * per_gpu * args.steps / (ms_total * 1e-3)

    # ---- per-kernel CUDA-event times (extra steps, events around every launch) ----------------------------
    prof_steps = min(args.steps, 3)
    net.profile_enable(True)
    for _ in range(prof_steps):
        net.forward_device(x, y, per_gpu, stream)
    torch.cuda.synchronize()
    prof = net.profile_read()
    net.profile_enable(False)

    # ---- end to end through the host-buffer call (what cuda::net_cuda::launch_forward makes) --------------
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((per_gpu, n_in), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hy = torch.empty((per_gpu, n_out), dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            net.forward_into(hx, hy)
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            net.forward_into(hx, hy)  # synchronous: returns when the logits are in host memory
        torch.cuda.synchronize()
        dt_sync = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        # the same calls, two in flight (netcuda_submit / netcuda_wait): step i+1's H2D overlaps step i's kernels.  Every step
        # still copies its own inputs up and its own logits down inside the timed region, which ends when the last logits landed.
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
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dt_sync, op=dist.ReduceOp.MAX)
        e2e = {"value": world * per_gpu * e2e_steps / float(dt.item()), "unit": "images/sec",
               "h2d_bytes_per_step": world * per_gpu * n_in * 4, "d2h_bytes_per_step": world * per_gpu * n_out * 4,
               "steps": e2e_steps, "api": "netcuda_submit / netcuda_wait, two calls in flight (pinned host buffers; H2D of call i+1 overlaps the kernels of call i)",
               "synchronous_value": world * per_gpu * e2e_steps / float(dt_sync.item()),
               "synchronous_api": "netcuda_forward, what net_cuda::launch_forward calls (blocking; 2-slot staged H2D inside the call)"}
        assert bool((hy2 == hy).all())
        parity_probe = float((hy.to(dev) - y).abs().max().item())  # same kernels, same inputs: must be identical
        e2e["max_abs_diff_vs_device_path"] = parity_probe
    t_c = time.perf_counter()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    clocks = sampler.window(t_a, t_b)
    sampler.stop()
    peaks = load_peaks()
    flops_per_image = net.flops_per_sample
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k in GEMM_LABELS)
    gemm_flops = sum(v["flops"] for k, v in prof.items() if k in GEMM_LABELS)
    gemm_launches = sum(v["launches"] for k, v in prof.items() if k in GEMM_LABELS)
    all_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak = peaks["sustained"]  # the GEMMs are timed inside a long step, under the power cap
    per_kernel = {k: {"launches": v["launches"] // prof_steps, "ms_per_step": round(v["ms"] / prof_steps, 4),
                      "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] > 0 and v["ms"] > 0 else None,
                      "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                  for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    roofline = {
        "kernel": "gemm_tn_tcgen05_kernel<bf16> (all GEMM launches of a step)", "bound": "tensor",
        "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
        "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
        "frac_of_burst_peak": round(achieved / peaks["burst"], 4),
        "flops_per_launch": gemm_flops / max(gemm_launches, 1), "ms_per_launch": gemm_ms / max(gemm_launches, 1),
        "algorithmic_bytes_per_launch": sum(v["bytes"] for k, v in prof.items() if k in GEMM_LABELS) / max(gemm_launches, 1),
        "share_of_step": round(gemm_ms / all_ms, 4) if all_ms > 0 else None,
        "traffic": ncu_traffic_per_launch(prof, prof_steps, min(max_batch, per_gpu)),
        "whole_step_tflops": round(value / world * flops_per_image / 1e12, 2),
        "whole_step_frac_of_burst_peak": round(value / world * flops_per_image / 1e12 / peaks["burst"], 4),
        "per_kernel": per_kernel,
    }

    cpu_baseline = None
    if not args.no_cpu_baseline:
        r = cpu_forward_rate(preset, budget_s=20.0, steps=1, warmup=0)
        cpu_baseline = {"value": r["value"], "unit": "images/sec", "cores": r["cores"], "kind": "port",
                        "sample": f"{r['images_per_step']} images of the workload, one pass, fp32 oracle (oracle/oracle_vit.c), "
                                  f"{r['seconds']:.1f} s on {os.cpu_count()} host cores"}

    line = {
        "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": args.workload, "net": preset, "images_per_gpu_per_step": per_gpu, "global_batch": world * per_gpu,
                   "pass_size": max_batch, "weights": "random-init, replicated per GPU", "sharding": f"batch x{world}, logits all-gather" if world > 1 else "single GPU",
                   "l2": f"inputs ({per_gpu * n_in * 4 >> 20} MiB) and per-pass activations exceed the 126 MB L2; no flush needed",
                   "flops_per_image": flops_per_image},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    if e2e is not None:
        line["e2e"] = e2e
    line["wall_s"] = {"timed_region": round(t_b - t_a, 3), "profile_and_e2e": round(t_c - t_b, 3)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
