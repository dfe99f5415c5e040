#!/bin/bash
# cluster streaming kernel: INT8 parity tests, 60 s of stress against the oracle, A/B probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "int8" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_pair.log
timeout 300 python tools/umma_stream_stress.py 60 2>&1 | tail -5
timeout 600 python tools/umma_pair_probe.py 2>&1 | tee gpurun_out/umma_pair_probe.log
