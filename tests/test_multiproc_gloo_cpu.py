"""world_size-2 and -3 `gloo` runs of the only exchange step of the multi-GPU path: each rank
holds its interleaved row tiles, rank 0 gathers and reassembles the frame.  The rows
are produced by the CPU oracle here; on the GPU box the same helper moves NCCL tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyoracle
import raymarching_engine_b200 as rm
from raymarching_engine_b200 import sharding
from conftest import scene_source

W, H, T = 64, 40, 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, full_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = np.load(full_path)
        rows = sharding.owned_rows(H, T, world, rank)
        local = torch.from_numpy(full[rows].copy())          # what this rank's present() would return
        got = sharding.gather_rows_to_rank0(local, H, T, dist)
        if rank == 0:
            np.save(out_path, got.numpy())
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_gather_reassembles_the_frame(world, tmp_path):
    src = scene_source("guide")
    s = rm.default_schema(src, rm.default_custom_settings(src), width=W, height=H)
    _, rgba = pyoracle.run_job("guide", s, nthreads=2)
    full_path, out_path = str(tmp_path / "full.npy"), str(tmp_path / "out.npy")
    np.save(full_path, rgba)
    mp.spawn(_worker, args=(world, _free_port(), full_path, out_path), nprocs=world, join=True)
    np.testing.assert_array_equal(np.load(out_path), rgba)
