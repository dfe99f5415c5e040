#!/bin/bash
# A/B on one box: the product library (warp index broadcast by a shuffle: tcgen05.mma / TMA operands in uniform registers) against a
# build with -DNETCUDA_NO_UNIFORM_WARP (lib_nouw: R2UR.BROADCAST waterfall before every MMA of the warp-indexed issuers).
# ViT-B bench step (per-kernel table), ViT-Tiny step, attention alone.
mkdir -p gpurun_out
for i in 1 2 3; do
for tag in "" _nouw; do
NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib$tag timeout 200 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']
print('lib$tag vit_b', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', {k: v['ms_per_step'] for k, v in pk.items() if v['ms_per_step'] > 0.5})"
done; done
for tag in "" _nouw ""; do
NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib$tag timeout 200 python bench.py --workload vit_tiny_16_224_b256 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']
print('lib$tag vit_tiny', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', {k: v['ms_per_step'] for k, v in pk.items() if v['ms_per_step'] > 0.05})"
done
for tag in "" _nouw; do echo "lib$tag attention alone:"; NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib$tag timeout 200 python tools/attn_sweep.py 2>&1 | grep "kernel  [0124]:"; done
