#!/bin/bash
# small-call host path: zero-copy input and whole-call graph A/B on config C1; MLP parity tests in both modes
mkdir -p gpurun_out
K="mlp or c1 or async or class or int8 or forward_device or pinned or graph"
timeout 600 python -m pytest tests/test_gpu_nets.py -k "$K" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_small_default.log 2>&1; echo "pytest default rc=$?"; tail -n 3 gpurun_out/pytest_small_default.log
NETCUDA_SMALL_CALL_GRAPH=1 timeout 600 python -m pytest tests/test_gpu_nets.py -k "$K" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_small_graph.log 2>&1; echo "pytest graph rc=$?"; tail -n 3 gpurun_out/pytest_small_graph.log
rm -f gpurun_out/c1_probe2.log
for rep in 1 2; do
echo "default (zc 16K, no graph)" | tee -a gpurun_out/c1_probe2.log
timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe2.log
echo "graph" | tee -a gpurun_out/c1_probe2.log
NETCUDA_SMALL_CALL_GRAPH=1 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe2.log
echo "no zero-copy input" | tee -a gpurun_out/c1_probe2.log
NETCUDA_SMALL_CALL_ZC_BYTES=0 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe2.log
echo "graph, no zero-copy input" | tee -a gpurun_out/c1_probe2.log
NETCUDA_SMALL_CALL_GRAPH=1 NETCUDA_SMALL_CALL_ZC_BYTES=0 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe2.log
done
