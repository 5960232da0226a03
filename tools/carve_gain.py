#!/usr/bin/env python3
"""Far-field pipeline on / off for one scene: ms per 1920x1080 preview frame through the public API (device-resident, no
readback), RMB_CARVE=1 vs 0 in child processes.   python tools/carve_gain.py menger-sponge"""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def child(name):
    import raymarching_engine_b200 as rm
    from conftest import scene_source
    src = scene_source(name)
    custom = rm.default_custom_settings(src)
    ctx = rm.load_render_job_context(device=0, specialize="always")
    prog = ctx.program_cache.get_program(src, None, custom)
    assert isinstance(prog, rm.Program), prog
    W, H = 1920, 1080
    L = rm._lib.lib

    def frame(k):
        s = rm.default_schema(src, custom, width=W, height=H, renderMode="preview", frameid=100 + k)
        s.camera.position = (0.3, 0.4, -3.0)
        fb = ctx.fbo.create(W, H, s.render.frameid)
        rm.upload_sample_uniforms(prog, s, (0.5, 1.0 / 3.0))
        assert L.rmb_render_sample(ctx.handle, prog.handle, fb.handle, 0, 0, W, H) == 0, ctx.last_error()
        assert L.rmb_present_device(ctx.handle, fb.handle, 1.0) == 0
        ctx.fbo.delete(W, H, s.render.frameid)
    for k in range(4):
        frame(k)
    ctx.sync()
    ctx.counters3(reset=True)
    t0 = time.perf_counter()
    n = 24
    for k in range(n):
        frame(4 + k)
    ctx.sync()
    dt = time.perf_counter() - t0
    ev, px, far = ctx.counters3()
    print(f"{name}: carve={prog.has_carve()} {1e3 * dt / n:.3f} ms/frame  {W * H * n / dt / 1e6:.1f} Mpx/s  evals/px {ev / max(px, 1):.2f}  far-field share {far / max(ev, 1):.3f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--child":
        child(sys.argv[1])
    else:
        for name in sys.argv[1:] or ["menger-sponge"]:
            for carve in ("1", "0"):
                subprocess.run([sys.executable, __file__, name, "--child"], env=dict(os.environ, RMB_CARVE=carve), check=True)
