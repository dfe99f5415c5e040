#!/bin/bash
# Round-2 ncu --set full captures of the two INT8 streaming kernels on config C5: the mma.sync kernel at one sample and the tcgen05
# kernel at 128 samples.  Each capture follows the same command without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD1="python tools/stream_once.py 1"
timeout 200 $CMD1 > gpurun_out/stream_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_i8_stream -s 4 -c 1 -f -o gpurun_out/prof_mlp_stream $CMD1 > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"; tail -n 2 gpurun_out/ncu_stream.log
CMD2="python tools/stream_once.py 128"
timeout 200 $CMD2 > gpurun_out/umma_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_i8_umma_stream -s 4 -c 1 -f -o gpurun_out/prof_mlp_umma_stream $CMD2 > gpurun_out/ncu_umma.log 2>&1
echo "ncu umma rc=$?"; tail -n 2 gpurun_out/ncu_umma.log
