#!/bin/bash
# Pass-size sweep of the ViT-B step (does L2 residency of the residual stream between kernels pay at small passes?)
# and a long run (sustained power-capped rate vs the default short timed region).
mkdir -p gpurun_out
for mb in 64 128 256 512 1024; do
  timeout 200 python bench.py --steps 20 --warmup 5 --max-batch $mb --no-cpu-baseline --no-e2e 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']
print('pass', d['config']['pass_size'], 'images/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],2) for k,v in pk.items() if v['ms_per_step']>0.5})"
done
timeout 280 python bench.py --steps 150 --warmup 20 --no-cpu-baseline --no-e2e 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('long run (150 steps): images/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'clk', d['clocks'])"
