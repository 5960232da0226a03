"""All GPUs of one box behind ONE render-job context, in one process (include/rmb.h "device groups").

`load_render_job_group(devices)` returns an object with the interface of `RenderJobContext` - `fbo.create/delete`,
`program_cache.get_program`, `render_sample`, `present` - so `do_render_job`, `run_job` and `make_presenter`
(renderer/RenderJobExecutor.tsx:77-341, index.tsx:25-59) drive it unchanged: the frame is split into interleaved row
tiles across the devices, `present` returns the assembled full frame.  Pure ctypes over
libraymarch_b200.so: no torch, no torch.distributed, no NCCL - the gather is peer stores from the display kernels into
member 0's frame, ordered by cross-device events (csrc/rmb_group.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from .executor import FramebufferInfo, Program, RenderJobContext, ShaderError
from .uniforms import UniformData

L = _lib.lib


class GroupProgram:
    is_group = True

    def __init__(self, context: "RenderJobGroup", handle: int):
        self.context, self.handle = context, handle

    def member(self, i: int) -> Program:
        return Program(self.context.member(i), L.rmb_group_program_member(self.handle, i))

    def source(self) -> str:
        return self.member(0).source()

    def has_carve(self) -> bool:
        return self.member(0).has_carve()


class GroupFramebufferInfo:
    def __init__(self, context: "RenderJobGroup", handle: int, width: int, height: int, frameid: int):
        self.context, self.handle = context, handle
        self.width, self.height, self.frameid = width, height, frameid
        self.local_rows = height          # what present() hands out: the whole frame

    def member(self, i: int) -> FramebufferInfo:
        return FramebufferInfo(self.context.member(i), L.rmb_group_fb_member(self.handle, i), self.width, self.height, self.frameid)

    def read(self, plane: str) -> np.ndarray:
        """the whole-frame accumulator plane, assembled on the host from the members' rows (tests)"""
        out = None
        for i in range(self.context.size):
            m = self.member(i)
            part = m.read(plane)
            if out is None:
                out = np.zeros((self.height,) + part.shape[1:], dtype=part.dtype)
            if m.local_rows:
                out[m.global_rows()] = part
        return out


class _GroupFbo:
    def __init__(self, g: "RenderJobGroup"):
        self._g = g

    def create(self, width: int, height: int, frameid: int) -> Optional[GroupFramebufferInfo]:
        h = L.rmb_group_fb_acquire(self._g.handle, int(width), int(height), int(frameid))
        return GroupFramebufferInfo(self._g, h, int(width), int(height), int(frameid)) if h else None

    def delete(self, width: int, height: int, frameid: int) -> None:
        L.rmb_group_fb_release(self._g.handle, int(width), int(height), int(frameid))


class _GroupProgramCache:
    def __init__(self, g: "RenderJobGroup"):
        self._g = g

    def get_program(self, scene_source: str, flavour: Optional[int] = None, spec: Optional[Dict[str, UniformData]] = None):
        flavour = self._g.flavour if flavour is None else flavour
        arr, n = _lib.make_spec_array(spec)
        out = C.c_void_p()
        etype = C.create_string_buffer(16)
        log = C.create_string_buffer(1 << 16)
        src = scene_source.encode()
        st = L.rmb_group_program_get(self._g.handle, src, len(src), flavour, arr, n, C.byref(out), etype, log, len(log))
        if st != _lib.RMB_OK:
            return ShaderError(etype.value.decode() or "general", log.value.decode())
        return GroupProgram(self._g, out.value)


class RenderJobGroup:
    """RenderJobContext (RenderJobExecutor.tsx:32-54) over n devices."""

    def __init__(self, handle: int, devices: Sequence[int], tile_rows: int, flavour: int, specialize):
        self.handle, self.devices, self.tile_rows = handle, list(devices), tile_rows
        self.size = len(self.devices)
        self.flavour, self.specialize = flavour, specialize
        self.fbo = _GroupFbo(self)
        self.program_cache = _GroupProgramCache(self)
        self._members: Dict[int, RenderJobContext] = {}
        self._pinned: Dict[tuple, tuple] = {}

    def member(self, i: int) -> RenderJobContext:
        """borrowed view of member context i (counters, timing, inspection); never close() it"""
        m = self._members.get(i)
        if m is None:
            m = RenderJobContext(L.rmb_group_ctx(self.handle, i), self.devices[i], i, self.size, self.tile_rows, self.flavour, self.specialize)
            self._members[i] = m
        return m

    def last_error(self) -> str:
        return (L.rmb_group_last_error(self.handle) or b"").decode()

    def sync(self) -> None:
        L.rmb_group_sync(self.handle)

    def render_sample(self, program: GroupProgram, fb: GroupFramebufferInfo, x: int, y: int, w: int, h: int) -> int:
        return L.rmb_group_render_sample(self.handle, program.handle, fb.handle, x, y, w, h)

    def _buf(self, key, shape, dtype) -> np.ndarray:
        k = (key, tuple(shape), np.dtype(dtype).str)
        b = self._pinned.get(k)
        if b is None:
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            ptr = L.rmb_host_alloc(max(nbytes, 1))
            if not ptr:
                raise MemoryError("rmb_host_alloc failed")
            arr = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(ptr))[:nbytes].view(dtype).reshape(shape)
            b = self._pinned[k] = (ptr, arr)
        return b[1]

    def present(self, fb: GroupFramebufferInfo, brightness: float, want_depth: bool = True, readback: bool = True):
        """the assembled frame: (rgba8[H, W, 4] uint8, depth[H, W] float32 | None); readback=False: display + assembly only"""
        if not readback:
            if L.rmb_group_present_device(self.handle, fb.handle, float(np.float32(brightness)), None) != _lib.RMB_OK:
                raise RuntimeError(self.last_error())
            return None, None
        rgba = self._buf("rgba", (fb.height, fb.width, 4), np.uint8)
        depth = self._buf("depth", (fb.height, fb.width), np.float32) if want_depth else None
        st = L.rmb_group_present(self.handle, fb.handle, float(np.float32(brightness)), rgba.ctypes.data_as(C.c_void_p),
                                 depth.ctypes.data_as(C.c_void_p) if want_depth else None)
        if st != _lib.RMB_OK:
            raise RuntimeError(self.last_error())
        return rgba, depth

    def present_async(self, fb, brightness: float, want_depth: bool = True, slot: int = 0):
        return self.present(fb, brightness, want_depth)          # the group's present blocks (rmb_group_present)

    def present_wait(self, fb) -> None:
        return None

    def close(self) -> None:
        if self.handle:
            L.rmb_group_destroy(self.handle)
            self.handle = None
            for m in self._members.values():
                m.handle = None
            for ptr, _a in self._pinned.values():
                L.rmb_host_free(ptr)
            self._pinned.clear()


def load_render_job_group(devices: Sequence[int], tile_rows: int = 16, flavour: int = _lib.FLAVOUR_EXACT,
                          specialize=None) -> Optional[RenderJobGroup]:
    """loadRenderJobContext over several GPUs of this box (devices may repeat: two members on one GPU).  None on
    failure, like load_render_job_context; group_error() explains."""
    if specialize is None:
        specialize = os.environ.get("RMB_SPECIALIZE", "auto")
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    h = L.rmb_group_create(arr, len(devices), int(tile_rows))
    if not h:
        return None
    return RenderJobGroup(h, devices, tile_rows, flavour, specialize)


def group_error() -> str:
    return (L.rmb_group_last_error(None) or b"").decode()
