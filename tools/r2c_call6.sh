#!/bin/bash
# 16 epilogue warps for short-K / single-column GEMMs: parity, then A/B on the ViT-Tiny step (bf16)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "gemm" -q -x --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -n 3 gpurun_out/pytest_gemm.log
timeout 600 python -m pytest tests/test_gpu_nets.py -k "vit" -q -x --timeout 400 -p no:cacheprovider > gpurun_out/pytest_vit.log 2>&1; echo "pytest vit rc=$?"; tail -n 3 gpurun_out/pytest_vit.log
B="--workload vit_tiny_16_224_b256 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
for rep in 1 2 3; do
for ew in 1 0; do
NETCUDA_GEMM_EW16=$ew timeout 300 python bench.py $B > gpurun_out/tiny_ew$ew.json 2> gpurun_out/tiny_ew$ew.err
python - <<P
import json
d=json.loads(open("gpurun_out/tiny_ew$ew.json").read().strip().splitlines()[-1])
pk=d["roofline"]["per_kernel"]
print("ew16=$ew", round(d["value"]), d["ms_per_step"], {k: pk[k]["ms_per_step"] for k in ("qkv","proj","fc1","fc2")})
P
done
done
