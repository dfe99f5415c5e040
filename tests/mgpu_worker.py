"""torchrun worker of tests/test_multi_gpu.py: one process per GPU, batch-sharded forward, NCCL all-gather of the logits,
compared BIT FOR BIT with the whole batch forwarded on one GPU ("1 vs 2/4/8 GPU bit-identical").

Every rank draws the same global batch (same seed), forwards its own contiguous slice (SURVEY.md s.8e partitioning) and takes part
in the gather; rank 0 then forwards the whole batch alone and asserts equality.  Prints one line `MGPU_OK world=<n>` on success.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
import netcuda as nc  # noqa: E402
from netcuda.sharding import gather_outputs, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    checks = []
    # (1) the golden ViT, ragged shards (37 images over `world` ranks), several internal passes per shard
    g = np.load(os.path.join(ROOT, "tests", "golden", "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    # (2) ViT-Tiny at full depth, 64 images per rank in one pass: the bench's kernels (CTA pairs, tcgen05 attention)
    tiny = nc.VIT_PRESETS["vit_tiny_16_224"]
    for name, c, flat, batch, max_batch in (("golden", cfg, g["flat"], 37, 8), ("vit_tiny", tiny, nc.vit_random_params(tiny, seed=1), 64 * world, 64)):
        net = nc.Net.vit(c, device=local, max_batch=max_batch)
        net.upload_vit(flat)
        gen = torch.Generator(device=dev).manual_seed(4321)
        x = torch.rand((batch, net.n_in), generator=gen, device=dev) * 2 - 1  # same global batch on every rank
        lo, hi = shard_bounds(batch, world, rank)
        y = torch.empty((hi - lo, net.n_out), device=dev)
        s = torch.cuda.current_stream()
        net.forward_device(x[lo:hi].contiguous(), y, hi - lo, s)
        full = gather_outputs(y, batch, world)
        torch.cuda.synchronize()
        if rank == 0:
            single = torch.empty((batch, net.n_out), device=dev)
            net.forward_device(x, single, batch, s)
            torch.cuda.synchronize()
            checks.append((name, bool(torch.equal(full, single)), float((full - single).abs().max())))
        net.close()
    # (3) config C5's INT8 net on a small batch: integers, so any difference at all is a bug
    from vit_presets import mlp_int8_params
    npl, n_ins = [512, 256, 64], 384
    wq, bq = mlp_int8_params(npl, n_ins, seed=3)
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, device=local)
    net.upload_mlp_i8(wq, bq)
    batch = 100 * world + 3
    gen = torch.Generator(device=dev).manual_seed(99)
    xq = torch.randint(-128, 128, (batch, n_ins), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)
    lo, hi = shard_bounds(batch, world, rank)
    yq = torch.empty((hi - lo, npl[-1]), dtype=torch.int32, device=dev)
    net.forward_device_i8(xq[lo:hi].contiguous(), yq, hi - lo, torch.cuda.current_stream())
    full = gather_outputs(yq, batch, world)
    torch.cuda.synchronize()
    if rank == 0:
        single = torch.empty((batch, npl[-1]), dtype=torch.int32, device=dev)
        net.forward_device_i8(xq, single, batch, torch.cuda.current_stream())
        torch.cuda.synchronize()
        checks.append(("int8_mlp", bool(torch.equal(full, single)), float((full - single).abs().max())))
    net.close()
    dist.barrier()
    if rank == 0:
        bad = [c for c in checks if not c[1]]
        print("checks:", checks, flush=True)
        if bad:
            raise SystemExit(f"sharded outputs differ from the single-GPU forward: {bad}")
        print(f"MGPU_OK world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
