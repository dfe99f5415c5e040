#!/usr/bin/env python
"""Turn the ncu artefacts under gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/make_profiles.py r1      # writes profiles/r1_*.{csv,md}

Inputs (from tools/gpu_ncu.sh on the GPU box): launches.csv (gpu__time_duration per launch), prof_*.ncu-rep (--set full).
"""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(P, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launch_list(fname="launches.csv", suffix="", what="`python bench.py --steps 1 --warmup 3 --batch 512 --no-cpu-baseline --no-e2e --no-configs` (weight upload + 4 warm-up + 1 timed + 1 profiled pass of 512 images)"):
    src = os.path.join(G, fname)
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = [["id", "kernel", "grid", "block", "duration_ns"]]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("nc::", "")
        ns = float(r[vi].replace(",", ""))
        out.append([r[0], name, r[gi], r[bi], f"{ns:.0f}"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    with open(os.path.join(P, f"{tag}_launches{suffix}.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launch_shares{suffix}.md"), "w") as f:
        f.write(f"# Launch list summary ({tag}{suffix})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over {what}.  "
                "Per-launch times are cold-cache and serialised: "
                "compare SHARES with `roofline.per_kernel` of the bench line, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / tot * 100:.1f} % |\n")


def full_reports():
    for rep in sorted(os.listdir(G)):
        if not rep.endswith(".ncu-rep"):
            continue
        raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        name = rep[:-8]
        with open(os.path.join(P, f"{tag}_{name}.md"), "w") as f:
            f.write(f"# ncu --set full --clock-control none: {name} ({tag})\n\nOne column per captured launch (B200, one kernel at a time, cold clocks).\n\n")
            f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |\n|---|---|" + "---:|" * (len(rows) - 2) + "\n")
            f.write("| kernel | | " + " | ".join("`" + r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("nc::", "")[:60] + "`" for r in rows[2:]) + " |\n")
            for m in METRICS:
                if m in idx:
                    f.write(f"| {m} | {units[idx[m]]} | " + " | ".join(r[idx[m]] for r in rows[2:]) + " |\n")


def traffic_json():
    """DRAM bytes per captured GEMM launch (qkv, proj, fc1, fc2 of one encoder block) for bench.py's roofline.traffic."""
    import json
    rep = os.path.join(G, "prof_gemm.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    labels = ["qkv", "proj", "fc1", "fc2"]
    out = {"source": f"profiles/{tag}_prof_gemm.md (ncu --set full, 512-image pass of ViT-B/16-224)", "pass_images": 512, "per_launch_bytes": {}}
    for lab, r in zip(labels, rows[2:]):
        out["per_launch_bytes"][lab] = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    json.dump(out, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)


# (every artefact found under gpurun_out/ is summarised under the given tag: clear gpurun_out/ of older captures first)
launch_list()
launch_list("launches_tiny.csv", "_vit_tiny", "`python bench.py --steps 1 --warmup 3 --workload vit_tiny_16_224_b256 --no-cpu-baseline --no-e2e --no-configs` (ViT-Tiny/16-224, 256 images per pass)")
full_reports()
traffic_json()
print(sorted(os.listdir(P)))
