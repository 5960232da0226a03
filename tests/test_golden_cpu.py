"""The committed golden fixtures (tests/golden/*.npz, written by tests/golden/gen_golden.py from the CPU
oracle) still come out of the oracle bit for bit: guards the oracle - and the shared deterministic math
layer under it - against silent drift.  No GPU needed."""
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import gen_golden  # noqa: E402


@pytest.mark.parametrize("name", sorted(gen_golden.CASES))
def test_oracle_reproduces_golden(name):
    import raymarching_engine_b200 as rm
    scene, s = gen_golden.make_case(rm, name)
    got = gen_golden.render_oracle(scene, s)
    want = np.load(HERE / "golden" / f"{name}.npz")
    for k in ("rgba8", "color", "nd", "ad", "depth"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=f"{name}:{k}")


def test_every_fixture_has_a_case():
    files = {p.stem for p in (HERE / "golden").glob("*.npz")}
    assert files == set(gen_golden.CASES)
