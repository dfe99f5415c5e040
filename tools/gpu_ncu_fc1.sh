#!/bin/bash
mkdir -p gpurun_out
for v in 0 3; do
CMD="python bench.py --steps 1 --warmup 3 --batch 256 --max-batch 256 --no-cpu-baseline --no-e2e --gemm-variant $v"
timeout 200 $CMD > gpurun_out/plain_v$v.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 31 -c 1 -f -o gpurun_out/prof_fc1_v$v $CMD > gpurun_out/ncu_fc1_v$v.log 2>&1
echo "ncu v$v rc=$?"
done
