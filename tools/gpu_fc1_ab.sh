#!/bin/bash
# fc1 (GELU epilogue) A/B: packed-fp32 GELU vs scalar, clock64 timeline, one ncu --set full capture with source.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -n 3
echo "== packed"; timeout 120 python tools/gemm_epi_ab.py 2>&1 | tail -n 6
echo "== scalar"; NETCUDA_GELU_SCALAR=1 timeout 120 python tools/gemm_epi_ab.py 2>&1 | tail -n 6
timeout 120 python tools/gemm_timeline.py > gpurun_out/gemm_timeline.log 2>&1; echo "timeline rc=$?"
CMD="python tools/gemm_epi_ab.py"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 11 -c 1 -f -o gpurun_out/prof_fc1_x2 $CMD > gpurun_out/ncu_fc1_x2.log 2>&1
echo "ncu rc=$?"
