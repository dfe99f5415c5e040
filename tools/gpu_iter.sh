#!/bin/bash
# Iteration loop on the GPU: parity suite (fail fast), bench line, optional extra command.
# Everything runs under its own `timeout`: a deadlocked kernel must not eat the box's time limit.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 30 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; cat gpurun_out/bench_iter.json; tail -n 5 gpurun_out/bench_iter.err
if [ -n "$1" ]; then timeout 300 bash -c "$1"; fi
