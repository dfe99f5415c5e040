#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "whole-step frac", d["roofline"]["whole_step_frac_of_burst_peak"], "wall", d.get("wall_s"))
    print("e2e", json.dumps({k: v for k, v in d.get("e2e", {}).items() if "api" not in k}))
    for k, v in d["roofline"]["per_kernel"].items(): print("  ", k, v)
    for k, v in d.get("configs", {}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline", {}).get("whole_step_frac_of_burst_peak"), v.get("e2e"))
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
for i in 1 2 3; do
EXTRA="" run att0_$i NETCUDA_ATT_TC_VARIANT=0
EXTRA="" run att4_$i NETCUDA_ATT_TC_VARIANT=4
done
