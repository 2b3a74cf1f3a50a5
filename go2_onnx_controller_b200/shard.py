"""Data-parallel sharding of a batch of robots / rollouts over the GPUs of one box.

Rows are independent, so each rank (one process per GPU) owns a contiguous block of rows, holds its own copy of
the 189 KB of weights and runs the same kernels on its block: there is NO collective on the hot path.  The only
communication is the start/stop barrier of a measurement and, off the hot path, an optional gather of the
[B,12] actions (48 B/row) to every rank for a consumer that wants them in one place.
"""
from __future__ import annotations

import os
from typing import Callable, Tuple


def shard_rows(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block (row0, rows) of rank `rank`; the first total_rows % world_size ranks get one extra row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    if total_rows < 0:
        raise ValueError("negative row count")
    base, extra = divmod(total_rows, world_size)
    row0 = rank * base + min(rank, extra)
    return row0, base + (1 if rank < extra else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) as torchrun exports them; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def gather_actions(local_actions, total_rows: int, group=None):
    """Off-hot-path gather of every rank's [rows_r, out_dim] action block into [total_rows, out_dim] on every
    rank (torch tensors; NCCL on GPUs, gloo on CPU).  Blocks may differ by one row, so the exchange is padded to
    the largest block and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out_dim = local_actions.shape[1]
    max_rows = (total_rows + world - 1) // world
    padded = local_actions.new_zeros((max_rows, out_dim))
    padded[: local_actions.shape[0]] = local_actions
    gathered = local_actions.new_empty((world * max_rows, out_dim))
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for r in range(world):
        _, rows = shard_rows(total_rows, world, r)
        parts.append(gathered[r * max_rows: r * max_rows + rows])
    return torch.cat(parts, 0)


def run_sharded(total_rows: int, rank: int, world_size: int, fn: Callable[[int, int], object]):
    """Calls fn(row0, rows) for this rank's block and returns its result."""
    row0, rows = shard_rows(total_rows, world_size, rank)
    return fn(row0, rows)
