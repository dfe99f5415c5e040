#!/bin/bash
# 16 epilogue warps for tf32 GEMMs with K <= 256: parity, then A/B on the TF32 ViT-Tiny step
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -k "short_k or tf32" -q -x --timeout 150 -p no:cacheprovider > gpurun_out/pytest_tf32_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 gpurun_out/pytest_tf32_k.log
timeout 300 python -m pytest tests/test_gpu_nets.py -k "tf32" -q -x --timeout 250 -p no:cacheprovider > gpurun_out/pytest_tf32_n.log 2>&1; echo "pytest nets rc=$?"; tail -n 2 gpurun_out/pytest_tf32_n.log
B="--workload vit_tiny_16_224_b256_tf32 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
for rep in 1 2; do
for ew in 1 0; do
NETCUDA_GEMM_EW16=$ew timeout 200 python bench.py $B > gpurun_out/tiny_tf32_ew$ew.json 2> gpurun_out/tiny_tf32_ew$ew.err
python - <<P
import json
d=json.loads(open("gpurun_out/tiny_tf32_ew$ew.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("tf32 ew16=$ew", round(d["value"]), d["ms_per_step"], {k: pk[k]["ms_per_step"] for k in ("qkv","proj","fc1","fc2")})
P
done
done
