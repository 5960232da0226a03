"""PNG capture: replaces `canvas.toDataURL()` of the reference's "save image" button
(/root/reference/client/src/index.tsx:470-476).  The presented RGBA8 frame (row 0 = bottom, OpenGL
convention) is written top row first, 8-bit RGBA, non-interlaced; zlib + struct only."""
from __future__ import annotations

import base64
import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(rgba8: np.ndarray, compress_level: int = 6) -> bytes:
    """rgba8: (H, W, 4) uint8 with row 0 at the BOTTOM (what present() returns)."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected an (H, W, 4) uint8 array")
    h, w, _ = a.shape
    rows = a[::-1]                                            # the canvas' top row is the last GL row
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rows.reshape(h, w * 4)], axis=1).tobytes()   # filter type 0 per scanline
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            _chunk(b"IDAT", zlib.compress(raw, compress_level)) + _chunk(b"IEND", b""))


def save_png(path, rgba8: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(rgba8))


def to_data_url(rgba8: np.ndarray) -> str:
    """the string `canvas.toDataURL()` hands to the reference's download link (index.tsx:471)"""
    return "data:image/png;base64," + base64.b64encode(encode_png(rgba8)).decode()


def decode_png(data: bytes) -> np.ndarray:
    """Inverse of encode_png for the subset it writes (tests): returns (H, W, 4) uint8, row 0 = bottom."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG")
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if crc != zlib.crc32(tag + body) & 0xFFFFFFFF:
            raise ValueError("bad CRC")
        if tag == b"IHDR":
            w, h, depth, ctype, _c, _f, inter = struct.unpack(">IIBBBBB", body)
            if (depth, ctype, inter) != (8, 6, 0):
                raise ValueError("unsupported PNG flavour")
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 4)
    if raw[:, 0].any():
        raise ValueError("unsupported scanline filter")
    return raw[:, 1:].reshape(h, w, 4)[::-1].copy()
