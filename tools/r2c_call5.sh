#!/bin/bash
# half-warp LayerNorm for D <= 192: parity + A/B on the ViT-Tiny step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k layernorm -q -x --timeout 200 -p no:cacheprovider > gpurun_out/pytest_ln.log 2>&1; echo "pytest ln rc=$?"; tail -n 3 gpurun_out/pytest_ln.log
timeout 600 python -m pytest tests/test_gpu_nets.py -k "vit" -q -x --timeout 400 -p no:cacheprovider > gpurun_out/pytest_vit.log 2>&1; echo "pytest vit rc=$?"; tail -n 3 gpurun_out/pytest_vit.log
B="--workload vit_tiny_16_224_b256 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
for rep in 1 2 3; do
for hw in 1 0; do
NETCUDA_LN_HALFWARP=$hw timeout 300 python bench.py $B > gpurun_out/tiny_hw$hw.json 2> gpurun_out/tiny_hw$hw.err
python - <<P
import json
d=json.loads(open("gpurun_out/tiny_hw$hw.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("halfwarp=$hw", round(d["value"]), d["ms_per_step"], "LN", pk["layernorm"]["ms_per_step"], pk["layernorm"]["gbs"])
P
done
done
