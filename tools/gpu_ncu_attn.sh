#!/bin/bash
# ncu --set full capture of the tcgen05 attention kernel alone (ViT-B shapes, 256 images), after the same command ran without ncu.
mkdir -p gpurun_out
CMD="python tools/attn_timeline.py"
timeout 200 $CMD > gpurun_out/attn_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn2 $CMD > gpurun_out/ncu_attn2.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_attn2.log
ncu -i gpurun_out/prof_attn2.ncu-rep --page source --csv > gpurun_out/src_attn2.csv 2>/dev/null; ls -la gpurun_out/prof_attn2.ncu-rep gpurun_out/src_attn2.csv
