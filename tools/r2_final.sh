#!/bin/bash
# Round-end rehearsal on one GPU: the full -m gpu suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 6 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke done ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 3 gpurun_out/bench_full.err
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 3 gpurun_out/bench_ref.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "whole-step frac", d["roofline"]["whole_step_frac_of_burst_peak"], "gemm frac", d["roofline"]["frac"], "wall", d.get("wall_s"))
    print("e2e", json.dumps({k: v for k, v in d.get("e2e", {}).items() if "api" not in k}))
    print("cpu", d["cpu_baseline"])
    for k, v in d["roofline"]["per_kernel"].items(): print("  ", k, v)
    for k, v in d.get("configs", {}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline", {}).get("whole_step_frac_of_burst_peak"), v.get("e2e"))
    print(json.dumps({k: (v["us_per_forward"], v["roofline"]["frac"]) for k, v in d["configs"]["C5"]["sweep"].items()}))
    print(json.dumps(d["configs"]["C1"]["precisions"]))
    r = json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
    print("reference:", r["value"], r["cpu_baseline"], {k: {kk: vv.get("value") for kk, vv in v.items()} for k, v in r.get("configs", {}).items()})
except Exception as e:
    print("parse failed", e)
PY
