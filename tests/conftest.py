import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "oracle") not in sys.path:
    sys.path.insert(0, str(ROOT / "oracle"))


# The suite pins the specialisation policy of do_render_job to "always" (custom uniforms baked from the first job on:
# the variant bench.py measures); the "auto" default and the dynamic variant have their own tests.
os.environ.setdefault("RMB_SPECIALIZE", "always")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import __graft_entry__ as g
    lib = ROOT / "raymarching_engine_b200" / "libraymarch_b200.so"
    orc = ROOT / "oracle" / "liboracle.so"
    if not lib.exists() or not orc.exists():
        g.build()


_ensure_built()

SCENES = ["guide", "fractal1", "menger-sponge", "tree", "smooth-tree", "rotation-fractal", "sphere-grid",
          "inline-default", "mandelbulb"]


def scene_source(name: str) -> str:
    """scenes/ holds the workloads bench.py and smoke() render; the reference's other example scenes are
    parity-test fixtures only and live under tests/fixtures/scenes/ (see scenes/README.md)."""
    for d in (ROOT / "scenes", ROOT / "tests" / "fixtures" / "scenes"):
        p = d / f"{name}.glsl"
        if p.exists():
            return p.read_text()
    raise FileNotFoundError(name)


@pytest.fixture(scope="session")
def ctx():
    import raymarching_engine_b200 as rm
    c = rm.load_render_job_context(device=0)
    if c is None:
        pytest.fail("no CUDA context: " + rm.context_error())
    yield c
    c.close()
