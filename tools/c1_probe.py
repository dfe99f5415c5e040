#!/usr/bin/env python
"""Config C1 (MLP 784-128-64-10) latencies alone: bench.py's bench_c1 without the CPU baselines.
   NETCUDA_SMALL_CALL=0 python tools/c1_probe.py   # the two-stream host pipeline for small calls as well (A/B)"""
import json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

import netcuda as nc  # (bench.py put vit-fpga_b200/ on sys.path)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
out = bench.bench_c1(nc, torch, dev, 0, 1, False)
print(json.dumps({"small_call": os.environ.get("NETCUDA_SMALL_CALL", "1"), "precisions": out["precisions"]}))
