#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/two_lane_probe.py > gpurun_out/two_lane.log 2>&1; echo "two-lane rc=$?"; cat gpurun_out/two_lane.log | tail -n 12
CMD="python tools/gemm_epi_ab.py"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 25 -c 1 -f -o gpurun_out/prof_fc1_x2 $CMD > gpurun_out/ncu_fc1_x2.log 2>&1
echo "ncu rc=$?"
