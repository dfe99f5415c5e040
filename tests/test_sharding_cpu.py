"""CPU tests of the N > 1 host path: batch sharding + output gather over torch.distributed (gloo, world_size 2).
The per-rank "forward" is the CPU oracle here -- the collective plumbing is what is under test; on the GPU box the same
functions run over NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest

from conftest import c1_net


def test_shard_bounds_cover_the_batch_exactly():
    from netcuda.sharding import shard_bounds

    for batch in (0, 1, 2, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            per = -(-batch // world) if batch else 0
            assert all(hi - lo <= per for lo, hi in spans)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, result_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "vit-fpga_b200"))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch
    import torch.distributed as dist
    from netcuda.sharding import gather_outputs, shard_bounds
    from oracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    npl, n_ins = [128, 64, 10], 784
    w, b = o.rand_init(1, 784 * 128 + 128 * 64 + 64 * 10, sum(npl))
    x = np.random.default_rng(1234).uniform(-1, 1, (batch, n_ins)).astype(np.float32)  # every rank draws the same global batch
    lo, hi = shard_bounds(batch, world, rank)
    local = o.mlp_forward(x[lo:hi], w, b, npl, n_ins, threads=1) if hi > lo else np.zeros((0, 10), np.float32)
    full = gather_outputs(torch.from_numpy(local), batch, world)
    np.save(os.path.join(result_dir, f"rank{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [64, 5, 1])
def test_sharded_forward_equals_unsharded_world2(oracle, tmp_path, batch):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    npl, n_ins, w, b = c1_net(oracle)
    x = np.random.default_rng(1234).uniform(-1, 1, (batch, n_ins)).astype(np.float32)
    want = oracle.mlp_forward(x, w, b, npl, n_ins)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        np.testing.assert_array_equal(got, want)  # same kernels per sample -> bit-identical, on every rank
