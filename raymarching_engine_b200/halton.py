"""Halton low-discrepancy sequence.  Mirrors /root/reference/client/src/util/Halton.tsx:1-19
(an endless generator; JavaScript numbers are doubles, as are Python floats)."""
from __future__ import annotations

from typing import Iterator


def halton(b: int) -> Iterator[float]:
    n, d = 0, 1
    while True:
        x = d - n
        if x == 1:
            n = 1
            d *= b
        else:
            y = d
            while x <= y:
                y /= b
            n = (b + 1) * y - x
        yield n / d
