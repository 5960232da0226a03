"""Uniform value records (pure Python, no native library): `UniformData = {type: "f"|"i"|"ui", count: 1..4, data}` and
the `u` constructors of /root/reference/client/src/renderer/Uniforms.tsx:1-32.  Split from uniforms.py so that the
host-only modules (schema.py, params.py) can be loaded without libraymarch_b200.so - bench.py's `--impl reference`
arm does exactly that."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class UniformData:
    type: str            # "f" | "i" | "ui"      Uniforms.tsx:8
    count: int           # 1..4                  Uniforms.tsx:1-5
    data: tuple

    def __post_init__(self):
        if self.type not in ("f", "i", "ui"):
            raise ValueError(f"bad uniform type {self.type!r}")
        if not (1 <= self.count <= 4) or len(self.data) != self.count:
            raise ValueError("uniform count must be 1..4 and match len(data)")


class u:  # noqa: N801  (name follows Uniforms.tsx:11 `export namespace u`)
    @staticmethod
    def float(x) -> UniformData:
        return UniformData("f", 1, (x,))

    @staticmethod
    def vec2(x, y) -> UniformData:
        return UniformData("f", 2, (x, y))

    @staticmethod
    def vec3(x, y, z) -> UniformData:
        return UniformData("f", 3, (x, y, z))

    @staticmethod
    def vec4(x, y, z, w) -> UniformData:
        return UniformData("f", 4, (x, y, z, w))

    @staticmethod
    def int(x) -> UniformData:
        return UniformData("i", 1, (x,))
