"""Scene parameter annotations: `uniform <type> <name>;` followed by `//@key=value` comments.

Mirrors the scanner of /root/reference/client/src/settings/shader-editor/CustomShaderParamParser.tsx:8-209
(with util/StringStream.tsx:11-40) and the value plumbing of settings/CustomSettings.tsx:148-173:
the `@default` values are what the reference renders a scene with until the user moves a slider.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

from .uniform_types import UniformData

# Validate.tsx:84-85
UNIFORM_VARIABLE_RE = re.compile(r"^(u?int|float|[iu]?vec[234])\s+[a-zA-Z_][a-zA-Z_0-9]*")
_KVP_RE = re.compile(r'^@\w+\s*\=\s*("[^"]*?"|\S+)')
_WS_RE = re.compile(r"^\s")


@dataclass
class CustomShaderParam:                       # Validate.tsx:59-74
    type: str = "f"                            # "f" | "i" | "ui"
    quantity: int = 1
    formats: List[str] = field(default_factory=lambda: ["numerical"])
    name: str = ""
    tooltip: Optional[str] = None
    internalName: str = ""
    min: Optional[float] = None
    max: Optional[float] = None
    step: Optional[float] = None
    sensitivity: Optional[float] = None
    scale: Optional[str] = None
    defaultValue: List[float] = field(default_factory=lambda: [0, 0, 0, 0])
    success: bool = True


@dataclass
class CustomShaderParamError:                  # Validate.tsx:76-81
    reason: str
    start: int
    end: int
    success: bool = False


class _Stream:                                  # util/StringStream.tsx:11-40
    def __init__(self, s: str):
        self.s, self.p = s, 0

    def done(self) -> bool:
        return self.p >= len(self.s)

    def pos(self) -> int:
        return self.p

    def match(self, pattern, no_consume: bool = False) -> Optional[str]:
        rest = self.s[self.p:]
        if isinstance(pattern, re.Pattern):
            m = pattern.search(rest)     # all patterns used are ^-anchored
            if m and m.group(0):
                if not no_consume:
                    self.p += len(m.group(0))
                return m.group(0)
            return None
        if rest.startswith(pattern):
            if not no_consume:
                self.p += len(pattern)
            return pattern
        return None

    def next(self, n: int) -> str:
        self.p += n
        return self.s[self.p - n:self.p]


def _js_number(s: str) -> float:
    """JavaScript `Number(string)` for the literals that occur in annotations."""
    t = s.strip()
    if t == "":
        return 0.0
    try:
        if re.fullmatch(r"[+-]?0[xX][0-9a-fA-F]+", t):
            return float(int(t, 16))
        if re.fullmatch(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|Infinity)", t):
            return float(t.replace("Infinity", "inf"))
    except ValueError:
        pass
    return math.nan


def get_custom_shader_params(src: str) -> List[Union[CustomShaderParam, CustomShaderParamError]]:
    """getCustomShaderParams (CustomShaderParamParser.tsx:8-209), same state machine."""
    stream = _Stream(src)
    out: List[Union[CustomShaderParam, CustomShaderParamError]] = []
    in_comment: Union[bool, str] = False
    cur = CustomShaderParam()
    state = 0
    first = True

    def flush():
        nonlocal cur, state, first
        state = 1
        if not first:
            out.append(cur)
            cur = CustomShaderParam()
        first = False
        stream.match(_WS_RE)

    while not stream.done():
        if stream.match("//") and not in_comment:            # :71-75
            in_comment = "line"
            continue
        if stream.match("/*") and not in_comment:            # :76-79
            in_comment = "block"
            continue
        if stream.match("*/") and in_comment == "block":     # :80-83
            in_comment = False
            continue
        if stream.match("\n", True) and in_comment == "line":  # :84-87 (newline is not consumed)
            in_comment = False
            continue
        if in_comment:                                        # :90-169
            kvp = stream.match(_KVP_RE)
            if kvp:
                parts = [e.strip() for e in kvp.split("=")]
                raw_key, raw_value = parts[0], parts[1] if len(parts) > 1 else ""
                key = raw_key[1:]
                value = raw_value[1:-1] if raw_value[:1] == '"' else raw_value
                if key in ("min", "max", "step", "sensitivity"):
                    num = _js_number(value)
                    if math.isnan(num):
                        out.append(CustomShaderParamError(f"Expected property '{key}' to be a number.", stream.pos() - len(value), stream.pos()))
                    setattr(cur, key, num)
                elif key == "scale":
                    if value == "log":
                        cur.scale = "log"
                elif key == "name":
                    cur.name = value
                elif key == "tooltip":
                    cur.tooltip = value
                elif key == "format":
                    cur.formats = []
                    for fmt in value.split("/"):
                        if fmt in ("numerical", "position", "color", "checkbox"):
                            if fmt not in cur.formats:
                                cur.formats.append(fmt)
                        else:
                            out.append(CustomShaderParamError(
                                f"Unknown input format '{fmt}'. Accepted values are \"numerical\", \"position\", \"color\", and \"checkbox\"",
                                stream.pos() - len(value), stream.pos()))
                elif key == "default":
                    vals = value.split(",")
                    if len(vals) != cur.quantity:
                        out.append(CustomShaderParamError(
                            f"This variable requires {cur.quantity} default values, but {len(vals)} were supplied. "
                            "Note that you need quotes if a value contains spaces.", stream.pos() - len(value), stream.pos()))
                    cur.defaultValue = [_js_number(v) for v in vals]
                continue
            stream.next(1)
        else:                                                 # :171-204
            if state == 1:
                stream.match(_WS_RE)
                decl = stream.match(UNIFORM_VARIABLE_RE)
                if decl:
                    typename, var_name = [e.strip() for e in re.split(r"\s+", decl)][:2]
                    if not typename or not var_name:
                        continue
                    cur.quantity = 1
                    cur.type = "f"
                    if typename[0] == "u":
                        cur.type = "ui"
                    if typename[0] == "i":
                        cur.type = "i"
                    if "vec" in typename:
                        cur.quantity = int(typename[-1])
                    cur.name = var_name
                    cur.internalName = var_name
                else:
                    state = 0
                continue
            if stream.match("uniform"):
                flush()
                continue
            stream.next(1)
    flush()
    return out


def default_custom_settings(src: str) -> Dict[str, UniformData]:
    """Uniform values a scene renders with before any user input: the `@default` of each
    successfully parsed parameter (settings/CustomSettings.tsx:148-173 feeds these into
    `schema.customShaderParameters`, index.tsx:127)."""
    result: Dict[str, UniformData] = {}
    for p in get_custom_shader_params(src):
        if not p.success or not p.internalName:
            continue
        vals = list(p.defaultValue)[: p.quantity]
        while len(vals) < p.quantity:
            vals.append(0)
        if p.type == "f":
            data = tuple(float(v) for v in vals)
        else:
            data = tuple(int(v) if not math.isnan(v) else 0 for v in vals)
        result[p.internalName] = UniformData(p.type, p.quantity, data)
    return result
