"""Uniform value records.  Mirrors /root/reference/client/src/renderer/Uniforms.tsx:1-46
(`UniformData = {type: "f"|"i"|"ui", count: 1..4, data}`, namespace `u`, `setUniforms`)."""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

from . import _lib
from .uniform_types import UniformData, u  # noqa: F401  (re-exported)


_TYPE = {"f": (_lib.UNIFORM_F, C.c_float), "i": (_lib.UNIFORM_I, C.c_int32), "ui": (_lib.UNIFORM_UI, C.c_uint32)}


def _fns(program):
    """(set, set_array, matrix4) of the single-device ABI, or their rmb_group_* twins for a group program"""
    L = _lib.lib
    if getattr(program, "is_group", False):
        return L.rmb_group_uniform_set, L.rmb_group_uniform_set_array, L.rmb_group_uniform_matrix4
    return L.rmb_uniform_set, L.rmb_uniform_set_array, L.rmb_uniform_matrix4


def set_uniforms(program, uniforms: Mapping[str, UniformData]) -> None:
    """`setUniforms(gl, program, uniforms)` (Uniforms.tsx:34-46): gl.uniform{count}{type}v per entry.
    Unknown names are ignored like a null uniform location."""
    for name, s in uniforms.items():
        code, ct = _TYPE[s.type]
        buf = (ct * s.count)(*[ct(v).value for v in s.data])
        st = _fns(program)[0](program.handle, name.encode(), code, s.count, C.cast(buf, C.c_void_p))
        if st != _lib.RMB_OK:
            raise RuntimeError(f"uniform {name}: {program.context.last_error()}")


def set_uniform_array(program, name: str, components: int, values: Sequence[float]) -> None:
    """gl.uniform1fv / gl.uniform3fv on an array uniform (RenderJobExecutor.tsx:268-291)."""
    n = len(values) // components
    if n == 0:
        return
    buf = (C.c_float * (n * components))(*[float(v) for v in values[: n * components]])
    _fns(program)[1](program.handle, name.encode(), _lib.UNIFORM_F, components, n, C.cast(buf, C.c_void_p))


def set_uniform_matrix4(program, name: str, m16: Sequence[float]) -> None:
    """gl.uniformMatrix4fv(loc, false, m) (RenderJobExecutor.tsx:293-297): column-major."""
    buf = (C.c_float * 16)(*[float(v) for v in m16])
    _fns(program)[2](program.handle, name.encode(), buf)
