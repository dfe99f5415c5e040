#!/bin/bash
# ncu captures of the hot kernels on a short bench invocation (one 256-image pass per step).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 29 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 7 -c 1 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
