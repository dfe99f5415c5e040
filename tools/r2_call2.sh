#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -5
NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib_dbg timeout 200 python tools/attn_timeline.py 2 3 13 2>&1 | grep -v "^warp  [57]\|^warp 10"
timeout 200 python tools/attn_sweep.py
