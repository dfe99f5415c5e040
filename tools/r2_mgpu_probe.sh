#!/bin/bash
nproc
timeout 300 python tools/class_multi_gpu_probe.py 2048 2>&1 | tail -4
NETCUDA_DEVICES=2 NETCUDA_SHARD_DEBUG=1 NETCUDA_HOST_TRACE=1 timeout 200 python - <<'PY' 2>&1 | tail -40
import os, sys
import numpy as np
sys.path.insert(0, "vit-fpga_b200")
import netcuda as nc
cfg = nc.VIT_PRESETS["vit_base_16_224"]
flat = nc.vit_random_params(cfg, seed=0)
x = np.random.default_rng(0).uniform(-1, 1, (2048, 3 * 224 * 224)).astype(np.float32)
net = nc.HostNet.vit(cfg, flat, max_batch=512)
dt, y = net.time_launch_forward(x, reps=1)
print(2048 / dt, "images/s")
PY
