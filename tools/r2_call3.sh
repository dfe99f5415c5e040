#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -4
timeout 200 python tools/attn_sweep.py
