"""Multi-GPU bit-identity on hardware: a torchrun-launched run, one process per GPU over NCCL,
whose gathered logits must equal the single-GPU forward of the same global batch bit for bit.

The driver's GPU test box has one GPU: there the test skips.  Set NETCUDA_REQUIRE_GPUS=<n> (e.g. under `gpurun --gpus 2`) to make a
missing GPU a FAILURE instead of a skip -- the form used for this repository's own 2-GPU verification (profiles/r2_multi_gpu.md).
The host-side sharding arithmetic is covered on CPU (gloo, world size 2) by tests/test_sharding_cpu.py.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_forward_equals_single_gpu_bit_for_bit(torch_cuda):
    want = int(os.environ.get("NETCUDA_REQUIRE_GPUS", "0"))
    have = torch_cuda.cuda.device_count()
    world = want if want > 0 else min(have, 2)
    if have < max(world, 2):
        if want > 0:
            pytest.fail(f"NETCUDA_REQUIRE_GPUS={want} but only {have} CUDA device(s) are visible")
        pytest.skip("needs two GPUs (set NETCUDA_REQUIRE_GPUS=2 to fail instead of skipping)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"MGPU_OK world={world}" in r.stdout, r.stdout[-2000:]
