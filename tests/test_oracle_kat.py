"""Known-answer tests that pin the CPU oracle to values derivable from the reference source
(SURVEY.md 8c).  The reference ships no golden vectors, so these are the anchors."""
import ctypes as C
import math

import numpy as np
import pytest

import pyoracle
import raymarching_engine_b200 as rm
from conftest import scene_source

L = pyoracle.lib()


def test_halton_sequences():
    # Halton.tsx:1-19
    assert pyoracle.halton_seq(2, 10) == [1 / 2, 1 / 4, 3 / 4, 1 / 8, 5 / 8, 3 / 8, 7 / 8, 1 / 16, 9 / 16, 5 / 16]
    want3 = [1 / 3, 2 / 3, 1 / 9, 4 / 9, 7 / 9, 2 / 9, 5 / 9, 8 / 9, 1 / 27, 10 / 27]
    assert np.allclose(pyoracle.halton_seq(3, 10), want3, rtol=0, atol=1e-15)


def test_guide_sdf_known_answers():
    # guide.glsl:91-102 with its @default uniforms
    sdf = lambda x, y, z: L.orc_sdf(b"guide", None, 0, x, y, z)
    assert sdf(0, 0, 0) == 6.0
    assert sdf(0, 0, 6) == 0.0
    assert sdf(3, 3, 3) == pytest.approx(4.18535, abs=2e-5)
    assert sdf(0, 0, 10) == pytest.approx(-1.67e-4, abs=2e-6)     # interior point
    # explicit uniforms in declaration order give the same result as the built-in defaults
    cu = pyoracle.flatten_custom("guide", rm.default_custom_settings(scene_source("guide")))
    assert L.orc_sdf(b"guide", cu.ctypes.data_as(C.c_void_p), cu.size, 3, 3, 3) == sdf(3, 3, 3)


def test_sdbox_known_answers():
    # raymarcher.frag:108-112
    assert L.orc_sdbox(0, 0, 0, 1, 1, 1) == -1.0
    assert L.orc_sdbox(2, 0, 0, 1, 1, 1) == 1.0
    assert L.orc_sdbox(2, 2, 0, 1, 1, 1) == pytest.approx(math.sqrt(2), rel=1e-6)


def test_other_scene_sdfs_against_closed_forms():
    # sphere-grid.glsl:46-49: spheres of radius 0.4 on the even lattice
    assert L.orc_sdf(b"sphere-grid", None, 0, 0, 0, 0) == pytest.approx(-0.4, abs=1e-6)
    assert L.orc_sdf(b"sphere-grid", None, 0, 1, 1, 1) == pytest.approx(math.sqrt(3) - 0.4, rel=1e-6)
    assert L.orc_sdf(b"sphere-grid", None, 0, 2, 0, 0) == pytest.approx(-0.4, abs=1e-6)
    # menger-sponge.glsl:7: far outside every cross the bounding box dominates
    assert L.orc_sdf(b"menger-sponge", None, 0, 3.0, -0.5, -0.5) == pytest.approx(3.0, rel=1e-6)
    # inline default (index.tsx:374-388): bounding sphere of radius 5 at the origin
    assert L.orc_sdf(b"inline-default", None, 0, 0, 0, 9) == pytest.approx(4.0, abs=1e-5)


def test_first_uniform_sample_seed():
    # uniformSample(): seed += 0.131223; noise seed = fract(randNoise.x + seed)   raymarcher.frag:91-94
    W, H, px, py = 1280, 720, 100, 50
    out = np.zeros(3, np.float32)
    L.orc_uniform_samples(0.5, 1 / 3, px, py, W, H, 3, out.ctypes.data_as(C.c_void_p))
    tx = np.float32((np.float32(px) + np.float32(0.5)) / np.float32(W)) * np.float32(1000)
    ty = np.float32((np.float32(py) + np.float32(0.5)) / np.float32(H)) * np.float32(1000)
    seed = np.float32(np.float32(0.5) + np.float32(0.131223))
    assert abs(float(seed) - 0.631223) < 1e-7          # fract(0.5 + 0.131223) = 0.631223
    assert out[0] == L.orc_gold_noise(float(tx), float(ty), float(seed))
    assert 0.0 <= out[0] < 1.0 and 0.0 <= out[1] < 1.0 and out[0] != out[1]


def test_gold_noise_definition():
    # fract(tan(distance(xy*PHI, xy)*seed)*xy.x), raymarcher.frag:46-49, re-evaluated step by step with
    # numpy float32 arithmetic (the fused dot product of distance() emulated in float64)
    f32 = np.float32
    PHI = f32(1.61803398874989484820459)
    for (x, y, sd) in [(10.5, 20.25, 0.631223), (400.0, 300.0, 0.25), (999.6, 0.7, 0.9), (78.515625, 69.44444, 0.63122296)]:
        xf, yf, sf = f32(x), f32(y), f32(sd)
        dx = f32(f32(xf * PHI) - xf)
        dy = f32(f32(yf * PHI) - yf)
        dot = f32(float(dy) * float(dy) + float(f32(dx * dx)))        # fma(dy, dy, dx*dx)
        dist = f32(math.sqrt(float(dot)))
        arg = f32(dist * sf)
        t = f32(L.orc_builtin(2, float(arg), 0.0))                    # the oracle's tan of that exact argument
        prod = f32(t * xf)
        want = f32(prod - f32(np.floor(prod)))
        got = L.orc_gold_noise(float(xf), float(yf), float(sf))
        assert got == want and 0.0 <= got < 1.0
        # and the oracle's tan agrees with float64 tan to within an ulp
        assert abs(float(t) - math.tan(float(arg))) <= abs(float(np.spacing(t)))


def test_f16_conversions_match_numpy():
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.normal(0, 1, 2000), rng.normal(0, 1e-6, 500), rng.uniform(-70000, 70000, 500),
                           [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.99e-8, np.inf, -np.inf]]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([L.orc_f32_to_f16(float(v)) for v in vals], np.uint16)
    np.testing.assert_array_equal(got, want)
    allh = np.arange(0, 65536, 7, dtype=np.uint16)
    back = np.array([L.orc_f16_to_f32(int(h)) for h in allh], np.float32)
    ref = allh.view(np.float16).astype(np.float32)
    ok = (back.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(back) & np.isnan(ref))
    assert ok.all()


def _default_uniforms(W, H, mode=1):
    s = rm.default_schema(scene_source("guide"), rm.default_custom_settings(scene_source("guide")), width=W, height=H,
                          renderMode="preview" if mode == 1 else "full")
    return s, pyoracle.uniforms_from_schema(s, (0.5, 1 / 3))


def test_default_view_statistics():
    # ~9.7 % of pixels hit the fractal; miss rays freeze near step 38; sky colour formula
    W, H = 320, 180
    s, U = _default_uniforms(W, H)
    acc = pyoracle.Accumulators(W, H)
    pyoracle.render_sample("guide", s.customShaderParameters, U, acc)
    hit = acc.depth < 1e3
    assert 0.090 < hit.mean() < 0.105
    steps = np.zeros((H, W), np.int32)
    L.orc_preview_exit_steps(b"guide", None, 0, C.byref(U), W, H, steps.ctypes.data_as(C.c_void_p))
    miss_steps = steps[~hit]
    assert 36 <= np.median(miss_steps) <= 40
    assert (steps < 128).mean() > 0.99
    # a sky pixel near the top: colour = exposure * 2 * (0.7, 0.8, 1.0) * max(dir.y, 0.2), alpha 0 (additive preview)
    tr = np.zeros(17, np.float32)
    L.orc_trace_pixel(b"guide", None, 0, C.byref(U), W, H, 10, 170, tr.ctypes.data_as(C.c_void_p))
    d = tr[3:6].astype(np.float64)
    end = tr[9:12].astype(np.float64)
    dy = max(end[1] / np.linalg.norm(end), 0.2)
    want = 0.5 * 2.0 * np.array([0.7, 0.8, 1.0]) * dy
    np.testing.assert_allclose(acc.color[170, 10, :3], want, rtol=2e-6)
    assert acc.color[170, 10, 3] == 0.0
    assert abs(np.linalg.norm(d) - 1.0) < 1e-6


def test_display_pass_gamma_and_rounding():
    # display.frag:54 with kernelSize 0: byte = round(255 * (c * brightness)^(1/2.2)), alpha 255
    W, H = 8, 4
    acc = pyoracle.Accumulators(W, H)
    vals = np.linspace(0, 1.2, W * H).reshape(H, W).astype(np.float32)
    acc.color[..., 0] = vals
    acc.color[..., 1] = vals * 0.5
    acc.color[..., 2] = 0.25
    for brightness in (1.0, 0.5):
        out = pyoracle.display(acc, brightness)
        want = np.floor(np.clip((acc.color[..., :3].astype(np.float64) * brightness) ** (1 / 2.2), 0, 1) * 255 + 0.5)
        assert np.abs(out[..., :3].astype(np.int64) - want.astype(np.int64)).max() <= 1    # float32 vs float64 ties
        assert (out[..., 3] == 255).all()


def test_display_blur_is_normalised_and_wraps():
    # a constant image stays constant under the DoF blur; REPEAT wrap pulls the opposite edge in
    W, H = 16, 12
    acc = pyoracle.Accumulators(W, H)
    acc.color[..., :3] = 0.5
    acc.nd[..., 3] = np.float16(0.02).view(np.uint16)      # kernelSize = 0.02*200 = 4
    out = pyoracle.display(acc, 1.0)
    assert out[..., 0].min() == out[..., 0].max() == int(math.floor(255 * 0.5 ** (1 / 2.2) + 0.5))
    acc.color[:, 0, :3] = 4.0                               # bright left column bleeds into the right edge
    out = pyoracle.display(acc, 1.0)
    assert out[5, W - 1, 0] > out[5, W // 2, 0]


def test_scissor_quirk_and_accumulation():
    # additive accumulation over two samples equals the sum; scissor limits the written region
    W, H = 48, 27
    s, U = _default_uniforms(W, H)
    a = pyoracle.Accumulators(W, H)
    pyoracle.render_sample("guide", s.customShaderParameters, U, a)
    first = a.color.copy()
    pyoracle.render_sample("guide", s.customShaderParameters, U, a)
    np.testing.assert_array_equal(a.color, first + first)   # same randNoise -> same sample, exact doubling
    b = pyoracle.Accumulators(W, H)
    pyoracle.render_sample("guide", s.customShaderParameters, U, b, scissor=(10, 5, 20, 8))
    mask = np.zeros((H, W), bool)
    mask[5:13, 10:30] = True
    assert (b.color[~mask] == 0).all() and (b.color[mask][:, :3].sum(axis=1) > 0).all()
    np.testing.assert_array_equal(b.color[mask], first[mask])


def test_full_mode_writes_aux_accumulators():
    W, H = 32, 18
    s = rm.default_schema(scene_source("guide"), rm.default_custom_settings(scene_source("guide")), width=W, height=H, renderMode="full")
    s.lights = [rm.default_light()]
    U = pyoracle.uniforms_from_schema(s, (0.5, 1 / 3))
    acc = pyoracle.Accumulators(W, H)
    pyoracle.render_sample("guide", s.customShaderParameters, U, acc)
    assert (acc.color[..., 3] == 1.0).all()                 # full mode adds alpha 1 (raymarcher.frag:386)
    depth16 = acc.ad[..., 3].view(np.float16).astype(np.float32)
    assert np.isfinite(acc.depth).all() and (acc.depth >= 1e-5).all()
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(depth16, acc.depth.astype(np.float16).astype(np.float32))   # RGBA16F storage of the fp32 depth
    assert not np.isnan(acc.color).any()
