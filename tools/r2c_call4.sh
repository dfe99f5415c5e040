#!/bin/bash
# 8-deep ring in the ordered fp32 latency kernel; staging ramp that ends on a pass boundary (e2e through launch_forward)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k fp32_ordered -q -x --timeout 200 -p no:cacheprovider > gpurun_out/pytest_fp32_small.log 2>&1; echo "pytest fp32_ordered rc=$?"; tail -n 3 gpurun_out/pytest_fp32_small.log
timeout 600 python -m pytest tests/test_gpu_nets.py -k "mlp or c1 or async or class or host_call or forward_device" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_nets_subset.log 2>&1; echo "pytest nets rc=$?"; tail -n 3 gpurun_out/pytest_nets_subset.log
timeout 200 python tools/c1_probe.py 2>&1 | tee gpurun_out/c1_probe4.log
for rep in 1 2; do
timeout 600 python bench.py --no-configs --no-cpu-baseline > gpurun_out/bench_e2e_$rep.json 2> gpurun_out/bench_e2e_$rep.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/bench_e2e_$rep.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pin", round(d["e2e"]["pin_inputs_value"]), "async", round(d["e2e"]["pinned_async_value"]), "blocking", round(d["e2e"]["pinned_blocking_value"]), d["clocks"])
P
done
