#!/bin/bash
# A/B on the GPU: parity suite, then the bench line with LayerNorm as its own kernel (default) and folded into the GEMMs.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 30 gpurun_out/pytest_gpu.log
for f in 0 1; do
NETCUDA_LN_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ln$f.json 2> gpurun_out/bench_ln$f.err; echo "bench ln_fused=$f rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ln$f.json"))
print(round(d["value"]), d["ms_per_step"], d["clocks"], {k:round(v["ms_per_step"],2) for k,v in d["roofline"]["per_kernel"].items()})
PY
tail -n 3 gpurun_out/bench_ln$f.err
done
