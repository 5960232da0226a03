"""GPU parity at BASELINE.json's own sizes: the CUDA path (C ABI, far-field pipeline ON - the shipped
default) against the CPU oracle, bit for bit, on the frames bench.py actually renders.

  config 2   1920x1080 default scene, preview AND full mode (1 light), orbit poses 0, 7, 40, 100
  config 3   3840x2160 default scene, preview, one pose
  config 4   Mandelbulb, reflectionIterationCounts = [512], 480x270, preview and full
  config 5   a 64-row scissor band of a 7680x4320 frame, 16 spp, Halton continuation of pose 3
             (sample s of pose k uses Halton index k*16 + s)

The oracle needs ~10 s (preview) / ~35 s (full) per 1080p frame on 16 host threads, so this file costs a few
minutes of CPU time on the GPU box.  Tolerance: BASELINE.json asks RGBA8 within 1/255 on >= 99.9 % of
pixels and hit depth within 1e-4 relative; the exact flavour is held to bit identity of all four
accumulator planes and of the RGBA8 bytes.
"""
import ctypes as C

import numpy as np
import pytest

import pyoracle
import raymarching_engine_b200 as rm
from raymarching_engine_b200 import _lib
from conftest import scene_source, ROOT

import sys
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (the camera path and the schema of the benchmark frames)

pytestmark = pytest.mark.gpu
L = _lib.lib

_FRAME = [500000]


def _canon(a):
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).copy()
    bits[np.isnan(a)] = 0x7fc00000
    return bits


def _bench_schema(scene, W, H, mode, pose, counts=None, spp=1):
    src = scene_source(scene)
    _FRAME[0] += 1
    return bench.make_schema(rm, src, rm.default_custom_settings(src), W, H, mode, pose, _FRAME[0], scene, counts, spp)


def _compare_frame(ctx, scene, schema):
    rm.reset_halton()
    fb = ctx.fbo.create(schema.render.width, schema.render.height, schema.render.frameid)
    got = rm.run_job(schema, ctx)
    assert got["success"], got["why"]
    rgba8, depth = got["rgba8"].copy(), got["depth"].copy()
    planes = {p: fb.read(p) for p in ("color", "normalAndDofRadius", "albedoAndDepth", "depth")}
    acc, want = pyoracle.run_job(scene, schema)
    # the north_star tolerance first (so that a failure reports how far off it is), then bit identity
    diff = np.abs(rgba8.astype(np.int32) - want.astype(np.int32)).max(axis=2)
    assert float((diff <= 1).mean()) >= 0.999
    rel = np.abs(depth - acc.depth) / np.maximum(np.abs(acc.depth), 1e-30)
    ok = np.isfinite(acc.depth)
    assert float(np.nanmax(np.where(ok, rel, 0.0))) <= 1e-4
    np.testing.assert_array_equal(rgba8, want)
    np.testing.assert_array_equal(_canon(planes["color"]), _canon(acc.color))
    np.testing.assert_array_equal(planes["normalAndDofRadius"], acc.nd)
    np.testing.assert_array_equal(planes["albedoAndDepth"], acc.ad)
    np.testing.assert_array_equal(_canon(planes["depth"]), _canon(acc.depth))
    np.testing.assert_array_equal(_canon(depth), _canon(acc.depth))


def _assert_far_field_on(ctx, scene, schema):
    prog = ctx.program_cache.get_program(schema.sdfShaderSource, None, dict(schema.customShaderParameters))
    assert isinstance(prog, rm.Program)
    assert prog.has_carve(), "the far-field pipeline is expected to be ON for the default scene"


@pytest.mark.parametrize("pose", [0, 7, 40, 100])
def test_config2_1080p_preview_vs_oracle(ctx, pose):
    s = _bench_schema("guide", 1920, 1080, "preview", pose)
    _assert_far_field_on(ctx, "guide", s)
    ctx.counters3(reset=True)
    _compare_frame(ctx, "guide", s)
    evals, px, far = ctx.counters3(reset=True)
    assert px == 1920 * 1080 and far > 0.5 * evals       # the far-field shortcut really ran


@pytest.mark.parametrize("pose", [0, 7, 40, 100])
def test_config2_1080p_full_one_light_vs_oracle(ctx, pose):
    _compare_frame(ctx, "guide", _bench_schema("guide", 1920, 1080, "full", pose))


def test_config3_4k_preview_vs_oracle(ctx):
    _compare_frame(ctx, "guide", _bench_schema("guide", 3840, 2160, "preview", 40))


@pytest.mark.parametrize("mode", ["preview", "full"])
def test_config4_mandelbulb_512_steps_vs_oracle(ctx, mode):
    _compare_frame(ctx, "mandelbulb", _bench_schema("mandelbulb", 480, 270, mode, 5, counts=[512.0]))


def test_config5_8k_16spp_band_halton_continuation_vs_oracle(ctx):
    """64 rows of pose 3 of config 5: 7680x4320, 16 spp (exposure/16), samples 48..63 of the Halton
    sequence, driven through the C ABI with an explicit scissor (rmb_render_sample) because the
    reference's job loop has no band parameter; oracle: the same 16 samples under the same scissor."""
    W, H, spp, pose, y0, rows = 7680, 4320, 16, 3, 2128, 64
    s = _bench_schema("guide", W, H, "preview", pose, spp=spp)
    prog = ctx.program_cache.get_program(s.sdfShaderSource, None, dict(s.customShaderParameters))
    assert isinstance(prog, rm.Program)
    h2, h3 = pyoracle.halton_seq(2, (pose + 1) * spp), pyoracle.halton_seq(3, (pose + 1) * spp)
    import itertools
    assert list(itertools.islice(rm.halton(2), 3)) == h2[:3]
    fb = ctx.fbo.create(W, H, s.render.frameid)
    acc = pyoracle.Accumulators(W, H)
    for k in range(pose * spp, (pose + 1) * spp):
        noise = (h2[k], h3[k])
        rm.upload_sample_uniforms(prog, s, noise)
        assert L.rmb_render_sample(ctx.handle, prog.handle, fb.handle, 0, y0, W, rows) == 0, ctx.last_error()
        pyoracle.render_sample("guide", s.customShaderParameters, pyoracle.uniforms_from_schema(s, noise), acc, (0, y0, W, rows))
    rgba8, depth = ctx.present(fb, 1.0 / spp)
    rgba8, depth = rgba8[y0:y0 + rows].copy(), depth[y0:y0 + rows].copy()
    color = fb.read("color")[y0:y0 + rows]
    ctx.fbo.delete(W, H, s.render.frameid)
    band = pyoracle.Accumulators(W, rows)
    band.color[:], band.nd[:], band.ad[:], band.depth[:] = acc.color[y0:y0 + rows], acc.nd[y0:y0 + rows], acc.ad[y0:y0 + rows], acc.depth[y0:y0 + rows]
    want = pyoracle.display(band, 1.0 / spp)       # preview mode: kernelSize == 0, the display pass is per pixel
    np.testing.assert_array_equal(_canon(color), _canon(band.color))
    np.testing.assert_array_equal(_canon(depth), _canon(band.depth))
    np.testing.assert_array_equal(rgba8, want)
