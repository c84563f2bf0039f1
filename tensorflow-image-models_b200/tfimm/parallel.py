"""Data-parallel inference across the GPUs of one box: one process per GPU, batch sharded on axis 0,
weights replicated, ONE all-gather of the logits per forward.

The reference has no distribution code at all (no tf.distribute / NCCL / Horovod call site), so this module is
new surface, kept deliberately small: images are independent at inference (BatchNorm uses moving statistics,
LayerNorm is per token), so there is no data-path collective other than gathering the (B/R, classes) logits --
<= 1 MB per rank, latency-bound, which is why it is a plain ``all_gather_into_tensor`` (NCCL over NVLink/NVSwitch
on GPUs, gloo in the CPU tests) rather than a fused kernel.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of a global batch owned by ``rank``: contiguous, sizes differ by at most one."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(local: torch.Tensor, global_batch: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gathers per-rank logits (rows in rank order) into the full (global_batch, ...) tensor on every rank.
    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest shard first."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    n_local = local.shape[0]
    if global_batch is None or global_batch % world == 0:
        out = torch.empty((world * n_local, *local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    n_max = -(-global_batch // world)
    padded = torch.zeros((n_max, *local.shape[1:]), device=local.device, dtype=local.dtype)
    padded[:n_local] = local
    out = torch.empty((world * n_max, *local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(global_batch, r, world)
        pieces.append(out[r * n_max: r * n_max + (hi - lo)])
    return torch.cat(pieces, dim=0)


def data_parallel_forward(model, x: torch.Tensor) -> torch.Tensor:
    """Every rank passes the same global batch (or just calls with its own shard and ``gather_logits``):
    runs the local shard through ``model`` and returns the gathered logits for the whole batch."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model(x)
    local = model(shard_batch(x))
    return gather_logits(local, global_batch=x.shape[0])
