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
    Completion is ordered WITHOUT a collective (sync="flags", the default): 32-bit counters in a POSIX shared-memory
    segment every rank maps and registers (rmb_host_register), written and awaited in stream order
    (rmb_stream_write_u32 / rmb_stream_wait_geq_u32 = cuStreamWriteValue32 / cuStreamWaitValue32).  Every use of frame
    slot s is a generation g of that slot (all ranks see the same sequence of calls, so they count alike):
      complete():  rank r > 0 writes done[r][s] = g after its stores; rank 0's stream waits done[r][s] >= g for all r;
      aim()/scatter(): rank r > 0 waits consumed[s] >= g - 1 before it overwrites the slot; rank 0 writes consumed[s'] = g'
                   for every frame it completed earlier on the same context (all its work for them - the caller's reads
                   of the assembled frame included - is behind it in stream order; a slot last used on another context
                   is ordered through an event first, rmb_ctx_wait_ctx).
    No rank ever waits on the host, and rank r never waits for rank s != 0.  sync="nccl" keeps the earlier variant - a
    one-element all-reduce per frame on the contexts' streams - as the comparison arm and the fallback when the driver
    has no stream memory operations.  torch.distributed is used at construction (handle exchange) and teardown only.

    contexts: the RenderJobContext(s) of THIS rank; dist: an initialised torch.distributed (NCCL)."""

    _MAXS = 64          # frame slots the flag segment has room for

    def __init__(self, contexts, width: int, height: int, dist, slots: int = 4, blur: bool = False, sync: str = "flags"):
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
        # ---- completion flags
        self._index = {id(c): k for k, c in enumerate(self.contexts)}
        self._gen = [0] * slots                         # uses of every slot so far
        self._cur = {}                                  # context index -> (slot, generation) of its frame in flight
        self._pending = [[] for _ in self.contexts]     # rank 0: frames completed on a context, consumed[] not yet written
        self._last_ctx = {}                             # rank 0: slot -> context index of its last use
        self.sync, self._shm, self._flags_addr = "nccl", None, 0
        if sync == "flags" and slots <= self._MAXS:
            from multiprocessing import shared_memory
            nbytes = 4 * self._MAXS * (self.world + 1)
            box = [None]
            if self.rank == 0:
                self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
                self._shm.buf[:nbytes] = bytes(nbytes)
                box = [self._shm.name]
            dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                self._shm = shared_memory.SharedMemory(name=box[0])
            self._flags_addr = C.addressof(C.c_char.from_buffer(self._shm.buf))
            ok = L.rmb_host_register(self._flags_addr, nbytes) == _lib.RMB_OK
            registered = ok
            if ok:
                # probe the stream memory operations once (a write of 0 to one of this rank's own counters)
                ok = L.rmb_stream_write_u32(c0.handle, self._done(self.rank, 0), 0) == _lib.RMB_OK
                c0.sync()
            votes = [None] * self.world
            dist.all_gather_object(votes, bool(ok))
            if all(votes):
                self.sync = "flags"
            else:
                if registered:
                    L.rmb_host_unregister(self._flags_addr)
                self._close_shm()

    def _done(self, rank: int, slot: int) -> int:
        return self._flags_addr + 4 * (rank * self._MAXS + slot)

    def _consumed(self, slot: int) -> int:
        return self._flags_addr + 4 * (self.world * self._MAXS + slot)

    def _close_shm(self) -> None:
        if self._shm is not None:
            self._flags_addr = 0
            shm, self._shm = self._shm, None
            try:
                shm.close()
            except BufferError:
                pass                                  # a ctypes view is still alive: the segment goes with the process
            if self.rank == 0:
                try:
                    shm.unlink()
                except FileNotFoundError:
                    pass

    def _begin(self, context, slot: int) -> None:
        """flow control of the frame about to be stored into `slot` through `context`"""
        c = self._index[id(context)]
        s = slot % self.slots
        self._gen[s] += 1
        g = self._gen[s]
        self._cur[c] = (s, g)
        if self.sync != "flags":
            return
        L = _lib.lib

        def check(st, ctx=context):
            if st != _lib.RMB_OK:
                raise RuntimeError(ctx.last_error())
        if self.rank == 0:
            other = self._last_ctx.get(s)
            if other is not None and other != c:
                # the slot's previous frame was completed (and read) on ANOTHER context's stream of this rank
                if any(ps == s for ps, _pg in self._pending[other]):
                    # ... and not released yet: order this stream behind that one, release the slot here
                    check(L.rmb_ctx_wait_ctx(context.handle, self.contexts[other].handle))
                    check(L.rmb_stream_write_u32(context.handle, self._consumed(s), g - 1))
                    self._pending[other] = [(ps, pg) for ps, pg in self._pending[other] if ps != s]
                else:
                    # ... whose stream has released it (or will): this stream's stores wait for that release too
                    check(L.rmb_stream_wait_geq_u32(context.handle, self._consumed(s), g - 1))
            for ps, pg in self._pending[c]:
                check(L.rmb_stream_write_u32(context.handle, self._consumed(ps), pg))
            self._pending[c] = []
            self._last_ctx[s] = c
        elif g > 1:
            check(L.rmb_stream_wait_geq_u32(context.handle, self._consumed(s), g - 1))

    def aim(self, context, slot: int) -> None:
        """the next present of `context` also writes into frame buffer `slot` (no-blur flow)"""
        self._begin(context, slot)
        if _lib.lib.rmb_ctx_set_gather_target(context.handle, self.ptrs[slot % self.slots], self.nbytes) != _lib.RMB_OK:
            raise RuntimeError(context.last_error())

    def scatter(self, context, fb, slot: int) -> None:
        """blur flow: this rank's rows of the colour and normal+dofRadius accumulators -> rank 0's planes"""
        L = _lib.lib
        self._begin(context, slot)
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
        """after this rank's stores of the current frame of `context`: order "every rank has stored its rows" before
        whatever rank 0 enqueues next on that context (all ranks call it, in the same order per context)"""
        if self.sync == "flags":
            L = _lib.lib
            c = self._index[id(context)]
            sl, g = self._cur[c]
            if self.rank != 0:
                if L.rmb_stream_write_u32(context.handle, self._done(self.rank, sl), g) != _lib.RMB_OK:
                    raise RuntimeError(context.last_error())
            else:
                for r in range(1, self.world):
                    if L.rmb_stream_wait_geq_u32(context.handle, self._done(r, sl), g) != _lib.RMB_OK:
                        raise RuntimeError(context.last_error())
                self._pending[c].append((sl, g))
            return
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
        if self.sync == "flags" and self._flags_addr:
            L.rmb_host_unregister(self._flags_addr)
        self._close_shm()
        for p in self._owned:
            L.rmb_device_free(self.contexts[0].handle, p)
        self._opened, self._owned = [], []
