#!/bin/bash
# full parity suite + C5 probe with the uniform-warp build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 600 python tools/umma_pair_probe.py 2>&1 | tee gpurun_out/umma_pair_probe.log
NETCUDA_LIB_DIR=/root/repo/vit-fpga_b200/lib_dbg timeout 300 python tools/umma_pair_timeline.py 2>&1 | tee gpurun_out/umma_pair_timeline.log
