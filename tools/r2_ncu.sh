#!/bin/bash
# Round-2 ncu evidence for profiles/: launch lists of a bench step (ViT-B) and of a ViT-Tiny step, --set full captures of the GEMMs of one
# encoder block, the 16-softmax-warp attention kernel + LayerNorm, the key-blocked attention kernel (ViT-L/384), the kind::tf32 and
# kind::i8 GEMMs.  Each ncu run is preceded by the same command without ncu (B200_PROFILING.md).
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
CMDL="python bench.py $B --workload vit_large_16_384_b64 --batch 32"
$CMDL > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc_long -s 10 -c 1 -f -o gpurun_out/prof_attn_long $CMDL > gpurun_out/ncu_attn_long.log 2>&1
echo "ncu attn long rc=$?"
CMDT="python bench.py $B --workload vit_tiny_16_224_b256"
$CMDT > gpurun_out/plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tiny.csv $CMDT > gpurun_out/ncu_launches_tiny.log 2>&1
echo "ncu tiny launches rc=$?"
CMDF="python bench.py $B --workload vit_tiny_16_224_b256_tf32"
$CMDF > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 29 -c 4 -f -o gpurun_out/prof_gemm_tf32 $CMDF > gpurun_out/ncu_gemm_tf32.log 2>&1
echo "ncu tf32 rc=$?"
CMDI="python bench.py $B --workload mlp_8x4096_int8_b16384"
$CMDI > gpurun_out/plain7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 10 -c 2 -f -o gpurun_out/prof_gemm_i8 $CMDI > gpurun_out/ncu_gemm_i8.log 2>&1
echo "ncu i8 rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
tail -3 gpurun_out/plain7.log | cut -c1-400
