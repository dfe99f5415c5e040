#!/bin/bash
# Runs the GPU parity suites in isolated processes (a trapped kernel kills only its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n 25 gpurun_out/$name.log; }
run k_gemm_bf16 tests/test_gpu_kernels.py -m gpu -k "gemm_bf16"
run k_gemm_other tests/test_gpu_kernels.py -m gpu -k "gemm and not gemm_bf16"
run k_misc tests/test_gpu_kernels.py -m gpu -k "not gemm"
run nets_mlp tests/test_gpu_nets.py -m gpu -k "not vit and not c5"
run nets_c5 tests/test_gpu_nets.py -m gpu -k "c5"
run nets_vit tests/test_gpu_nets.py -m gpu -k "vit"
