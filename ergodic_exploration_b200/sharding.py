"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

The hot path shards by construction:
  * control(): instances are independent, so they are block-partitioned over
    the ranks, every rank keeps its instances' ut_/memory resident for the
    whole run, there is NO data-path collective inside a step, and the first
    twists are collected with one all_gather per step.
  * phi_k: the density grid is partitioned by row blocks, every rank contracts
    its rows (eb_phik_execute_raw_dev) and ONE all_reduce(sum) of the 32x32 raw
    block (which includes sum(Phi) at [0,0]) finishes the job.
Nothing here computes; it only decides who owns what and moves results.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of `total` units: the first total % world
    ranks get one extra unit.  Returns [lo, hi)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Gathers the ranks' row blocks (shard_bounds order) into the full
    (total, ...) tensor on every rank.  Uneven shards are padded to the largest
    shard so a single all_gather_into_tensor moves everything."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: local block has {local.shape[0]} rows, expected {sizes[rank]}")
    if world == 1:
        return local.clone()
    biggest = max(sizes)
    padded = local
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(s == biggest for s in sizes):
        return out
    return torch.cat([out[r * biggest: r * biggest + sizes[r]] for r in range(world)], dim=0)


def finish_phik(raw_local: torch.Tensor, nb: int, group=None):
    """raw_local: this rank's (32, 32) un-normalised partial contraction.
    all_reduce(sum), then phi_k[ky*nb + kx] = raw[ky, kx] / raw[0, 0]
    (target.cpp:87: the density is normalised to sum to one).  Returns
    (phi_k (nb*nb,), sum(Phi))."""
    raw = raw_local.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=group)
    total = raw[0, 0]
    return (raw[:nb, :nb] / total).reshape(-1), total


class ShardedErgodicControl:
    """B instances block-partitioned over the ranks of a process group; each
    rank drives an ErgodicControl for its own block on its own GPU."""

    def __init__(self, make_local, total_batch: int, group=None):
        """make_local(batch) -> ErgodicControl for `batch` instances on this rank's device"""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total = total_batch
        self.lo, self.hi = shard_bounds(total_batch, self.world, self.rank)
        self.local = make_local(self.hi - self.lo)

    def local_rows(self, full):
        return full[self.lo: self.hi]

    def control(self, grid, x_local: torch.Tensor, gather: bool = True):
        u0 = self.local.control(grid, x_local)
        if not gather or self.world == 1:
            return u0
        return all_gather_rows(u0, self.total, self.group)
