#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "tcgen05" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_pair.log
timeout 600 python tools/umma_pair_probe.py 2>&1 | tee gpurun_out/umma_pair_probe.log

NETCUDA_LIB_DIR=/root/repo/vit-fpga_b200/lib_dbg timeout 300 python tools/umma_pair_timeline.py 2>&1 | tee gpurun_out/umma_pair_timeline.log
