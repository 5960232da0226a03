"""Row-tile sharding of a frame across the GPUs of one box (BASELINE.json config 3).

Global tile t (tile_rows rows; the last tile may be partial) belongs to rank t % n_ranks.  Each rank
keeps only its own rows, in ascending global order.  The only exchange step is the gather of the
presented rows (RGBA8, optionally depth) to rank 0 - torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is plumbing here, the kernels never communicate.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _lib


def owned_rows(height: int, tile_rows: int, n_ranks: int, rank: int) -> np.ndarray:
    """global row indices owned by `rank`, ascending"""
    rows: List[int] = []
    n_tiles = -(-height // tile_rows)
    for t in range(rank, n_tiles, n_ranks):
        rows.extend(range(t * tile_rows, min((t + 1) * tile_rows, height)))
    return np.asarray(rows, dtype=np.int64)


def owned_rows_below(g: int, height: int, tile_rows: int, n_ranks: int, rank: int) -> int:
    """C implementation used by the library for scissor clipping (rmb_owned_rows_below)"""
    return _lib.lib.rmb_owned_rows_below(g, height, tile_rows, n_ranks, rank)


def max_local_rows(height: int, tile_rows: int, n_ranks: int) -> int:
    return max(len(owned_rows(height, tile_rows, n_ranks, r)) for r in range(n_ranks))


def gather_rows_to_rank0(local, height: int, tile_rows: int, dist, device=None):
    """Gathers every rank's local rows (torch tensor [local_rows, ...]) to rank 0 and reassembles the
    full frame [height, ...] there; returns None on the other ranks.  Ranks own different row counts,
    so the exchange uses equal-sized padded buffers."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    pad_rows = max_local_rows(height, tile_rows, world)
    send = torch.zeros((pad_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    recv: Optional[list] = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0)
    if rank != 0:
        return None
    full = torch.empty((height,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        rows = torch.as_tensor(owned_rows(height, tile_rows, world, r), device=local.device)
        full[rows] = recv[r][: len(rows)]
    return full
