#!/bin/bash
# Round 2, first GPU call: parity suite, attention variant sweep, GELU epilogue A/B, full bench line, step-level A/B runs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python tools/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/attn_sweep.log
timeout 200 python tools/gemm_epi_ab.py > gpurun_out/gemm_epi_ab.log 2>&1; echo "epi rc=$?"; cat gpurun_out/gemm_epi_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "whole-step frac", d["roofline"]["whole_step_frac_of_burst_peak"], "e2e", d.get("e2e", {}).get("value"),
          "pinned async", d.get("e2e", {}).get("pinned_async_value"), "wall", d.get("wall_s"))
    for k, v in d["roofline"]["per_kernel"].items(): print("  ", k, v)
    for k, v in d.get("configs", {}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline", {}).get("whole_step_frac_of_burst_peak"))
    print(json.dumps(d.get("configs", {}).get("C5", {}).get("sweep"), indent=0))
    print(json.dumps(d.get("configs", {}).get("C1", {}).get("precisions"), indent=0))
except Exception as e:
    print("parse failed", e)
PY
short="--steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-e2e"
run() { tag=$1; shift; timeout 200 env "$@" python bench.py $short $EXTRA > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    pk = d["roofline"]["per_kernel"]
    print(sys.argv[1], "value %.0f ms %.3f" % (d["value"], d["ms_per_step"]), {k: v["ms_per_step"] for k, v in pk.items() if v["ms_per_step"] > 0.2}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
EXTRA="" run base0 NETCUDA_ATT_TC_VARIANT=0
EXTRA="" run att2 NETCUDA_ATT_TC_VARIANT=2
EXTRA="" run att1 NETCUDA_ATT_TC_VARIANT=1
EXTRA="" run att22 NETCUDA_ATT_TC_VARIANT=22
EXTRA="" run att12 NETCUDA_ATT_TC_VARIANT=12
EXTRA="" run pdl NETCUDA_PDL=1
EXTRA="--max-batch 1024" run pass1024 NETCUDA_PDL=0
EXTRA="--max-batch 1024" run graph1024 NETCUDA_VIT_GRAPH_ROWS=100000000
EXTRA="--max-batch 1024" run graphpdl1024 NETCUDA_VIT_GRAPH_ROWS=100000000 NETCUDA_PDL=1
EXTRA="" run base0b NETCUDA_ATT_TC_VARIANT=0
