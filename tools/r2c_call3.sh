#!/bin/bash
# validation of the tree with the small-call path and the ordered fp32 latency kernel: parity suite, smoke, C1 graph A/B, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
rm -f gpurun_out/c1_probe3.log
for rep in 1 2; do
echo "mlp graphs from 2 layers (default)" | tee -a gpurun_out/c1_probe3.log
timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe3.log
echo "mlp graphs from 4 layers" | tee -a gpurun_out/c1_probe3.log
NETCUDA_MLP_GRAPH_MIN_LAYERS=4 timeout 200 python tools/c1_probe.py 2>&1 | tee -a gpurun_out/c1_probe3.log
done
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -c 300 gpurun_out/bench_ref.err
head -c 400 gpurun_out/bench_full.json
