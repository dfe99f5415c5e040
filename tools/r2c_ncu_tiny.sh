#!/bin/bash
# ncu --set full of the four GEMMs and the attention kernel of one encoder block inside a ViT-Tiny pass (256 images)
mkdir -p gpurun_out
T="python bench.py --steps 1 --warmup 3 --workload vit_tiny_16_224_b256 --no-cpu-baseline --no-e2e --no-configs"
timeout 200 $T > gpurun_out/tiny_plain3.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_tcgen05|attention_tc16" -s 26 -c 5 -f -o gpurun_out/prof_vit_tiny_block $T > gpurun_out/ncu_tiny_block.log 2>&1
echo "ncu tiny block rc=$?"; ls -la gpurun_out/*.ncu-rep
