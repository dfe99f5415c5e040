#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + --set full captures of the hot kernels.
# Each ncu run is preceded by the same command without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 512 --no-cpu-baseline --no-e2e"  # one pass of the bench's pass size (512 images)
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 29 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_tc|layernorm_kernel" -s 21 -c 2 -f -o gpurun_out/prof_attn_ln $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
