#!/bin/bash
# round-2c validation of the final tree: parity suite, smoke, both bench arms, ncu evidence of what changed (ViT-Tiny launch list,
# half-warp LayerNorm, ordered fp32 latency kernel).  Each ncu run is preceded by the same command without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_full.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -c 300 gpurun_out/bench_ref.err
T="python bench.py --steps 1 --warmup 3 --workload vit_tiny_16_224_b256 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 $T > gpurun_out/tiny_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_tiny.csv $T > gpurun_out/ncu_launches_tiny.log 2>&1
echo "ncu tiny launches rc=$?"
timeout 300 $T > gpurun_out/tiny_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_halfwarp -s 30 -c 1 -f -o gpurun_out/prof_layernorm_halfwarp $T > gpurun_out/ncu_ln.log 2>&1
echo "ncu ln rc=$?"
timeout 200 python tools/c1_once.py > gpurun_out/c1_once.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_fp32_ordered_small -s 12 -c 3 -f -o gpurun_out/prof_fp32_ordered_small python tools/c1_once.py > gpurun_out/ncu_fp32_small.log 2>&1
echo "ncu fp32 small rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
head -c 300 gpurun_out/bench_full.json
