"""Device groups (include/rmb.h rmb_group_*): all GPUs of one box behind one handle, through the C ABI alone - ctypes, no
torch, no NCCL.  The assembled frame of a group must equal the single-context frame bit for bit, in preview mode (display
kernels store into member 0's frame) and in full mode with depth of field (accumulator rows scattered to member 0, which
presents the assembled planes).  On a 1-GPU box the members share device 0 - the same code path without peer traffic, so
display.cu's gather store and the scatter kernel are exercised everywhere; with >= 2 GPUs the worker also uses two devices."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.environ["RM_ROOT"])
import raymarching_engine_b200 as rm
L = rm._lib.lib
ndev = int(sys.argv[1])
src = open(os.path.join(os.environ["RM_ROOT"], "scenes", "guide.glsl")).read()
custom = rm.default_custom_settings(src)
one = rm.load_render_job_context(device=0, specialize="always")
ok = True
serial = 0
for devices in ([0, 0], [0, 0, 0]) + (([0, 1],) if ndev >= 2 else ()) + (([0, 1, 2, 3],) if ndev >= 4 else ()):
    g = rm.load_render_job_group(devices, tile_rows=16, specialize="always")
    assert g is not None, rm.group_error()
    for frame, (mode, W, H) in enumerate([("preview", 200, 120), ("full", 200, 120), ("full", 101, 77)]):
        s = rm.default_schema(src, custom, width=W, height=H, renderMode=mode, frameid=10 + frame)
        if mode == "full":
            s.lights = [rm.default_light()]
            s.dof.amount = 0.05          # visible blur: the display pass reads neighbour tiles
            s.render.samplesPerPixel = 2
        rm.reset_halton()
        got = rm.run_job(s, g)           # the reference's doRenderJob loop, unchanged, over the group
        assert got["success"], got["why"]
        got = dict(got, rgba8=got["rgba8"].copy(), depth=got["depth"].copy())
        rm.reset_halton()
        serial += 1
        s.render.frameid = 100 + serial      # a fresh accumulator set for the comparison frame
        want = rm.run_job(s, one)
        same = bool(np.array_equal(got["rgba8"], want["rgba8"])) and bool(np.array_equal(got["depth"].view(np.uint32), want["depth"].view(np.uint32)))
        print(f"devices {devices} {mode} {W}x{H}: identical to the single-context frame = {same}", flush=True)
        ok = ok and same and got["rgba8"].shape == (H, W, 4)
    g.close()
one.close()
assert "torch" not in sys.modules, "the group path must not need torch"
print("torch imported:", "torch" in sys.modules, flush=True)
sys.exit(0 if ok else 1)
"""


def _device_count() -> int:
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
    return sum(1 for l in out.stdout.splitlines() if l.startswith("GPU "))


def test_group_frames_equal_single_context_through_the_c_abi_only(tmp_path):
    script = tmp_path / "group_worker.py"
    script.write_text(WORKER)
    n = _device_count()
    p = subprocess.run([sys.executable, str(script), str(n)], env=dict(os.environ, RM_ROOT=str(ROOT), RMB_SPECIALIZE="always"),
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("identical to the single-context frame = True") == 3 * (2 + (n >= 2) + (n >= 4))
    assert "torch imported: False" in p.stdout


def test_group_errors_are_values():
    import ctypes as C
    import raymarching_engine_b200 as rm
    L = rm._lib.lib
    assert rm.load_render_job_group([0, 99]) is None and "99" in rm.group_error()
    g = rm.load_render_job_group([0, 0])
    try:
        bad = g.program_cache.get_program("float sdf(vec3 p) { return nonsense; }")
        assert isinstance(bad, rm.ShaderError) and bad.type == "fragment" and "nonsense" in bad.infoLog
        assert L.rmb_group_size(g.handle) == 2 and L.rmb_group_ctx(g.handle, 2) is None
        assert g.fbo.create(0, 10, 1) is None
    finally:
        g.close()
