#!/bin/bash
# One GPU call: full parity suite, the bench line, an ncu launch list and an ncu --set full capture of the GEMM.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_vitb.json 2> gpurun_out/bench_vitb.err; echo "bench rc=$?"; cat gpurun_out/bench_vitb.json; tail -n 5 gpurun_out/bench_vitb.err
CMD="python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05 -s 29 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|layernorm_kernel" -s 14 -c 2 -f -o gpurun_out/prof_attn_ln $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out
