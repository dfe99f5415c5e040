#!/bin/bash
# Split-K pair streaming kernel: parity tests of the INT8 streaming kernels, then the A/B timing probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "int8" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_pair.log
timeout 600 python tools/umma_pair_probe.py 2>&1 | tee gpurun_out/umma_pair_probe.log
