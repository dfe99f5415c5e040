#!/bin/bash
# Multi-GPU verification on one box (gpurun --gpus N): bit-identity tests (fail, not skip), the bench line at N GPUs, and the C++ class
# driving N GPUs in one process.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
NETCUDA_REQUIRE_GPUS=$N timeout 900 python -m pytest tests/test_multi_gpu.py "tests/test_gpu_nets.py::test_cpp_class_shards_over_gpus" -m gpu -q -x --timeout 800 -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    keys = ("value", "ms_per_step", "n_gpus", "sharded_equals_single", "sharded_max_abs_diff_vs_single", "per_rank_ms_per_step", "all_gather_ms", "forward_only_ms_per_step_max_rank", "strong_scaling")
    print(json.dumps({k: d.get(k) for k in keys}, indent=1))
    print("e2e", json.dumps({k: v for k, v in d.get("e2e", {}).items() if "api" not in k}))
    print("C4", {k: d.get("configs", {}).get("C4", {}).get(k) for k in ("value", "ms_per_step", "n_gpus")})
    print("wall", d.get("wall_s"))
except Exception as e:
    print("parse failed", e)
PY
timeout 900 python tools/class_multi_gpu_probe.py $((N * 1024)) 2>&1 | tail -5
