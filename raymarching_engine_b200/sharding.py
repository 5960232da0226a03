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


class FusedTileGather:
    """Assembles row-tile sharded frames on rank 0 WITHOUT a pixel-moving collective: rank 0 owns `slots`
    full-frame buffers, exports them with CUDA IPC, and every rank stores its rows straight into them over
    NVLink.  Two flows:
      * no display blur (preview mode): aim() -> present: the display kernel itself writes RGBA8 at the
        global rows (rmb_ctx_set_gather_target);
      * blur=True (full mode with depth of field, whose display pass reads +-16 neighbour rows): scatter()
        copies the colour and normal+dofRadius accumulators into full-frame planes, rank 0 runs the
        display pass over them with display_assembled().
    The only collective is a one-element all-reduce per frame on the contexts' streams (complete()) that
    orders "every rank has stored its rows".

    contexts: the RenderJobContext(s) of THIS rank; dist: an initialised torch.distributed (NCCL)."""

    def __init__(self, contexts, width: int, height: int, dist, slots: int = 4, blur: bool = False):
        import ctypes as C
        import torch
        self.dist, self.contexts = dist, list(contexts)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.width, self.height, self.slots, self.blur = width, height, slots, blur
        self.nbytes = width * height * 4
        L = _lib.lib
        c0 = self.contexts[0]
        self._owned, self._opened = [], []
        planes = {"rgba8": 4}
        if blur:
            planes.update({"color": 16, "nd": 8})
        handles = None
        if self.rank == 0:
            handles = {}
            for name, bpp in planes.items():
                handles[name] = []
                for _ in range(slots):
                    p = L.rmb_device_alloc(c0.handle, width * height * bpp)
                    if not p:
                        raise MemoryError("rmb_device_alloc failed")
                    self._owned.append(p)
                    buf = C.create_string_buffer(64)
                    if L.rmb_ipc_export(p, buf) != _lib.RMB_OK:
                        raise RuntimeError("rmb_ipc_export failed: " + (L.rmb_last_error(None) or b"").decode())
                    handles[name].append((p, buf.raw))
        box = [handles]
        dist.broadcast_object_list(box, src=0)
        handles = box[0]
        self.planes = {}
        for name in planes:
            if self.rank == 0:
                self.planes[name] = [p for p, _h in handles[name]]
            else:
                self.planes[name] = []
                for _p, h in handles[name]:
                    out = C.c_void_p()
                    if L.rmb_ipc_open(c0.handle, h, C.byref(out)) != _lib.RMB_OK:
                        raise RuntimeError("rmb_ipc_open failed: " + c0.last_error())
                    self._opened.append(out.value)
                    self.planes[name].append(out.value)
        self.ptrs = self.planes["rgba8"]
        self.flag = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", c0.device))
        self._streams = {id(c): torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", c.device)) for c in self.contexts}
        self._torch = torch

    def aim(self, context, slot: int) -> None:
        """the next present of `context` also writes into frame buffer `slot` (no-blur flow)"""
        if _lib.lib.rmb_ctx_set_gather_target(context.handle, self.ptrs[slot % self.slots], self.nbytes) != _lib.RMB_OK:
            raise RuntimeError(context.last_error())

    def scatter(self, context, fb, slot: int) -> None:
        """blur flow: this rank's rows of the colour and normal+dofRadius accumulators -> rank 0's planes"""
        L = _lib.lib
        for which, name in ((0, "color"), (1, "nd")):
            if L.rmb_fb_scatter_rows(context.handle, fb.handle, which, self.planes[name][slot % self.slots]) != _lib.RMB_OK:
                raise RuntimeError(context.last_error())

    def display_assembled(self, context, slot: int, brightness: float) -> None:
        """blur flow, rank 0, after complete(): display pass over the assembled planes -> frame buffer `slot`"""
        assert self.rank == 0
        s = slot % self.slots
        if _lib.lib.rmb_display_planes(context.handle, self.planes["color"][s], self.planes["nd"][s], self.ptrs[s], self.width,
                                       self.height, float(brightness)) != _lib.RMB_OK:
            raise RuntimeError(context.last_error())

    def complete(self, context) -> None:
        """enqueue the per-frame completion collective on `context`'s stream (all ranks, same order)"""
        with self._torch.cuda.stream(self._streams[id(context)]):
            self.dist.all_reduce(self.flag)

    def frame_tensor(self, slot: int):
        """rank 0 only: the assembled frame [H, W, 4] uint8 as a torch view of the device buffer"""
        assert self.rank == 0
        import ctypes as C
        torch = self._torch
        n = self.nbytes
        # wrap the raw allocation without copying: go through the CUDA array interface
        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.height, self.width, 4), "typestr": "|u1", "data": (self.ptrs[slot % self.slots], False), "version": 2}
        return torch.as_tensor(raw, device=torch.device("cuda", self.contexts[0].device))

    def close(self) -> None:
        L = _lib.lib
        for c in self.contexts:
            L.rmb_ctx_set_gather_target(c.handle, None, 0)
            c.sync()
        for p in self._opened:
            L.rmb_ipc_close(self.contexts[0].handle, p)
        self.dist.barrier()
        for p in self._owned:
            L.rmb_device_free(self.contexts[0].handle, p)
        self._opened, self._owned = [], []
