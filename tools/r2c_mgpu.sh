#!/bin/bash
# 2-GPU evidence of the final round-2c tree: bit-identity tests (fail, not skip) and both bench arms under torchrun
N=${1:-2}
mkdir -p gpurun_out
NETCUDA_REQUIRE_GPUS=$N timeout 600 python -m pytest tests/test_multi_gpu.py "tests/test_gpu_nets.py::test_cpp_class_shards_over_gpus" -m gpu -q -x --timeout 500 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
keys = ("value", "ms_per_step", "n_gpus", "sharded_equals_single", "per_rank_ms_per_step", "all_gather_ms", "strong_scaling")
print(json.dumps({k: d.get(k) for k in keys}))
print("e2e", json.dumps({k: v for k, v in d.get("e2e", {}).items() if "api" not in k}))
print("wall", d.get("wall_s"))
PY
