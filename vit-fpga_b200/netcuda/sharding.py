"""Batch sharding across the GPUs of one box: independent samples, replicated weights, one
process per GPU, no data-path collective; only the outputs are gathered (NCCL on GPUs, gloo in the CPU tests).

    GPU g of G forwards samples [g * ceil(B / G), min(B, (g + 1) * ceil(B / G)))

This module holds the host-side logic only; it never computes a forward pass itself.
"""
from __future__ import annotations


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of the global batch owned by `rank`."""
    if world <= 0 or not 0 <= rank < world or batch < 0:
        raise ValueError(f"bad shard request batch={batch} world={world} rank={rank}")
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def gather_outputs(local, batch: int, world: int, group=None):
    """All-gather the per-rank output rows [shard, n_out] into the global [batch, n_out] tensor (same on every rank).
    Shards may be ragged (the last ranks can hold fewer rows or none): they are padded to ceil(batch / world) rows for
    the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    per = -(-batch // world)
    n_out = local.shape[1]
    padded = local
    if local.shape[0] != per:
        padded = torch.zeros((per, n_out), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * per, n_out), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:batch]
