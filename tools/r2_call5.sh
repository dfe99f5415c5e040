#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py -m gpu -q -x -k "golden or cpp_class or tf32" --timeout 300 -p no:cacheprovider 2>&1 | tail -4
short="--steps 10 --warmup 4 --no-configs --no-cpu-baseline --no-e2e"
run() { tag=$1; shift; timeout 200 env "$@" python bench.py $short $EXTRA > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f ms %.3f launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed", e, open(f"gpurun_out/ab_{sys.argv[1]}.err").read()[-300:])
PY
}
G="NETCUDA_VIT_GRAPH_ROWS=100000000 NETCUDA_VIT_GRAPH_MULTIPASS=1"
EXTRA="" run p512
EXTRA="--max-batch 256" run p256 $G
EXTRA="--max-batch 128" run p128 $G
EXTRA="--max-batch 128" run p128pdl $G NETCUDA_PDL=1
EXTRA="--max-batch 96" run p96pdl $G NETCUDA_PDL=1
EXTRA="--max-batch 64" run p64 $G
EXTRA="--max-batch 64" run p64pdl $G NETCUDA_PDL=1
EXTRA="--max-batch 32" run p32pdl $G NETCUDA_PDL=1
EXTRA="--max-batch 64" run p64nograph NETCUDA_PDL=1
EXTRA="" run p512b
