#!/bin/bash
# fused fp32 MLP forward (one cluster launch): parity, then config C1 with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k fp32_ordered -q -x --timeout 200 -p no:cacheprovider > gpurun_out/pytest_fp32_small.log 2>&1; echo "pytest fp32_ordered rc=$?"; tail -n 3 gpurun_out/pytest_fp32_small.log
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_weight_files.py -k "not vit and not int8 and not c5" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_nets_subset.log 2>&1; echo "pytest nets rc=$?"; tail -n 5 gpurun_out/pytest_nets_subset.log
rm -f gpurun_out/c1_probe8.log
for rep in 1 2; do
echo "fused (default)" | tee -a gpurun_out/c1_probe8.log
timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe8.log
echo "one launch per layer" | tee -a gpurun_out/c1_probe8.log
NETCUDA_MLP_FUSED=0 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe8.log
done
