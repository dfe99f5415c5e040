#!/bin/bash
# latency-oriented ordered fp32 kernel + single-stream small-call host path: parity suite, then config C1 with and without the small-call path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k fp32_ordered -q -x --timeout 200 -p no:cacheprovider > gpurun_out/pytest_fp32_small.log 2>&1; echo "pytest fp32_ordered rc=$?"; tail -n 5 gpurun_out/pytest_fp32_small.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 200 python tools/c1_probe.py 2>&1 | tee gpurun_out/c1_probe.log
NETCUDA_SMALL_CALL=0 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe.log
