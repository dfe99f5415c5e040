#!/bin/bash
# ncu --set full capture of the INT8 weight-streaming kernel (C5, one sample), after the same command ran without ncu.
mkdir -p gpurun_out
CMD="python tools/stream_once.py 1"
timeout 200 $CMD > gpurun_out/stream_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_i8_stream -s 4 -c 1 -f -o gpurun_out/prof_mlp_stream $CMD > gpurun_out/ncu_stream.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_stream.log
