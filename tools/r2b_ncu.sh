#!/bin/bash
# ncu evidence of the final round-2 library (after the uniform-warp-index change touched every tcgen05 kernel and the INT8 streaming
# kernel became split-K clusters): launch list of a ViT-B bench step, --set full captures of the GEMMs of one encoder block, of the
# attention kernel + LayerNorm, and of the cluster streaming kernel at 64 samples (four CTAs per tile) and 128 samples (pairs).
# Each ncu run is preceded by the same command without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
CMD="python bench.py $B --batch 512"  # one pass of the bench's pass size (512 images)
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 29 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_tc|layernorm_kernel" -s 21 -c 2 -f -o gpurun_out/prof_attn_ln $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
for n in 64 128; do
CMDS="python tools/stream_once.py $n"
timeout 200 $CMDS > gpurun_out/cluster_plain_$n.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_i8_umma_cluster -s 4 -c 1 -f -o gpurun_out/prof_mlp_umma_cluster_b$n $CMDS > gpurun_out/ncu_cluster_$n.log 2>&1
echo "ncu cluster $n rc=$?"; tail -n 1 gpurun_out/ncu_cluster_$n.log
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
