#!/usr/bin/env python3
"""Generates tests/golden/*.npz with the CPU oracle (oracle/, the restatement of the reference shader:
the reference itself ships no golden images and cannot run here - DESIGN.md section 2).  The fixtures
pin the oracle against silent drift (tests/test_golden_cpu.py re-renders them on the CPU) and give the
GPU tests a target that does not need the oracle's arithmetic to be re-evaluated identically at test
time (tests/test_parity_gpu.py::test_golden_fixtures).

usage: python tests/golden/gen_golden.py        (rewrites every fixture; commit the result)"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

# name -> (scene, W, H, mode, lights, spp, extra schema tweaks)
CASES = {
    "guide_preview_64x36": ("guide", 64, 36, "preview", 0, 1, {}),
    "guide_preview_pose_80x45_spp2": ("guide", 80, 45, "preview", 0, 2, {"pose": 0.7}),
    "guide_full_1light_48x27": ("guide", 48, 27, "full", 1, 1, {}),
    "guide_full_2lights_fog_mix_40x24_spp2": ("guide", 40, 24, "full", 2, 2, {"fog": 0.05, "blend": "mix"}),
    "sphere-grid_full_1light_48x27": ("sphere-grid", 48, 27, "full", 1, 1, {}),
    "menger-sponge_preview_64x36": ("menger-sponge", 64, 36, "preview", 0, 1, {}),
    "mandelbulb_preview_48x27": ("mandelbulb", 48, 27, "preview", 0, 1, {}),
    "tree_preview_ortho_48x32": ("tree", 48, 32, "preview", 0, 1, {"ortho": 12.0}),
}


def make_case(rm, name):
    import math
    scene, W, H, mode, lights, spp, extra = CASES[name]
    src = next(d / f"{scene}.glsl" for d in (ROOT / "scenes", ROOT / "tests" / "fixtures" / "scenes") if (d / f"{scene}.glsl").exists()).read_text()
    s = rm.default_schema(src, rm.default_custom_settings(src), width=W, height=H, renderMode=mode, samplesPerPixel=spp, frameid=1)
    s.lights = [rm.default_light() for _ in range(lights)]
    if lights > 1:
        s.lights[1].position = (2.0, 3.0, 4.0)
        s.lights[1].size = 0.5
    if "fog" in extra:
        s.fogDensity = extra["fog"]
    if "blend" in extra:
        s.render.blendMode = extra["blend"]
    if "ortho" in extra:
        s.camera.mode = rm.Orthographic(extra["ortho"])
    if "pose" in extra:
        th = extra["pose"]
        s.camera.position = (10.0 * math.sin(th), 0.0, 10.0 - 10.0 * math.cos(th))
        c, sn = math.cos(-th), math.sin(-th)
        s.camera.rotation = (c, 0, -sn, 0, 0, 1, 0, 0, sn, 0, c, 0, 0, 0, 0, 1)
    return scene, s


def render_oracle(scene, schema):
    import pyoracle
    acc, rgba = pyoracle.run_job(scene, schema)
    return {"rgba8": rgba, "color": acc.color.view(np.uint32), "nd": acc.nd, "ad": acc.ad, "depth": acc.depth.view(np.uint32)}


def main():
    # schema helpers only (pure Python); no rendering goes through the product here
    import raymarching_engine_b200.schema as schema_mod
    import raymarching_engine_b200.params as params_mod

    class RM:  # the subset of the package surface make_case needs, without loading the CUDA library
        default_schema = staticmethod(schema_mod.default_schema)
        default_light = staticmethod(schema_mod.default_light)
        Orthographic = schema_mod.Orthographic
        default_custom_settings = staticmethod(params_mod.default_custom_settings)
    for name in CASES:
        scene, s = make_case(RM, name)
        out = render_oracle(scene, s)
        np.savez_compressed(HERE / f"{name}.npz", **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
