"""Independent cross-check of the oracle: a numpy float64 restatement of the same GLSL, written
without the shared math layer, must agree with the C++ oracle to float32 accuracy on everything
that is not chaotic (SDF values, transcendental functions, camera ray, march of a pixel).  This is
what stands in for the golden vectors the reference does not have."""
import ctypes as C
import math

import numpy as np
import pytest

import pyoracle
import raymarching_engine_b200 as rm
from conftest import scene_source

L = pyoracle.lib()


def glsl_mod(x, y):
    return x - y * np.floor(x / y)


def sdf_guide(p, iters=8.0, gsf=float(np.float32(0.33333333333)), center=(0, 0, 10), size=4.0):
    p = np.asarray(p, dtype=np.float64)
    m = 9999.9
    i = -1.0
    while i < iters:
        sf = gsf ** i
        d = np.abs(glsl_mod(p + 0.5 * sf, sf) - sf / 2.0) - sf / 3.0
        m = min(math.sqrt(float(d @ d)) - 0.21 * sf, m)
        i += 1.0
    return max(np.linalg.norm(p - np.asarray(center, float)) - size, -m)


def sd_box(p, b):
    q = np.abs(p) - b
    return np.linalg.norm(np.maximum(q, 0.0)) + min(max(q[0], max(q[1], q[2])), 0.0)


def sdf_menger(p, iters=8.0):
    p = np.asarray(p, float)
    m = sd_box(p + 0.5, np.array([0.5] * 3))
    i = 1.0
    third = float(np.float32(0.33333333333333))
    while i < iters:
        sf = third ** i
        g = glsl_mod(p, sf * 3.0) - sf * 1.5
        m = max(m, -min(min(sd_box(g, np.array([sf * 1.51, sf * 0.5, sf * 0.5])), sd_box(g, np.array([sf * 0.5, sf * 1.51, sf * 0.5]))),
                        sd_box(g, np.array([sf * 0.5, sf * 0.5, sf * 1.51]))))
        i += 1.0
    return m


def sdf_tree(p, iters=8.0, scale=float(np.float32(0.7)), angles=(2.9, -0.8, 0.4), offset=1.2):
    angles = [float(np.float32(a)) for a in angles]
    offset = float(np.float32(offset))
    t = np.asarray(p, float).copy()
    m = 9999.0
    i = 0.0
    base = np.array([1.0, float(np.float32(0.1)), float(np.float32(0.1))])
    while i < iters:
        csf = scale ** i
        m = min(m, sd_box(t * csf, base * csf))
        t = t / scale
        t = np.abs(t) - base * offset
        for (a, b), ang in zip(((0, 1), (1, 2), (0, 2)), angles):
            c, s = math.cos(ang), math.sin(ang)
            na, nb = t[a] * c - t[b] * s, t[a] * s + t[b] * c
            t[a], t[b] = na, nb
        i += 1.0
    return m


@pytest.mark.parametrize("name,fn", [("guide", sdf_guide), ("menger-sponge", sdf_menger), ("tree", sdf_tree)])
def test_sdf_matches_float64_restatement(name, fn):
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-3, 3, (150, 3)) + (np.array([0, 0, 10]) if name == "guide" else 0), rng.uniform(-40, 40, (50, 3))])
    worst = 0.0
    for p in pts.astype(np.float32):
        got = L.orc_sdf(name.encode(), None, 0, float(p[0]), float(p[1]), float(p[2]))
        want = fn(p.astype(np.float64))
        # fp32 evaluation of a fractal fold: errors scale with the magnitudes involved, and a point that
        # sits on a fold boundary may pick the neighbouring cell (same distance by symmetry)
        tol = 2e-4 * max(1.0, abs(want)) if name != "tree" else 2e-3 * max(1.0, abs(want))
        worst = max(worst, abs(got - want) / max(1.0, abs(want)))
        assert abs(got - want) <= tol, (name, p, got, want)
    assert worst < 2e-3


def test_transcendentals_within_one_ulp_of_float64():
    rng = np.random.default_rng(5)
    ops = {0: np.sin, 1: np.cos, 2: np.tan, 4: np.exp, 5: np.log, 6: np.exp2, 7: np.log2, 8: np.sqrt, 17: np.arcsin, 18: np.arccos,
           19: np.arctan, 21: np.sinh, 22: np.cosh, 23: np.tanh, 30: np.arcsinh, 31: np.arccosh, 32: np.arctanh}
    dom = {4: (-80, 80), 6: (-120, 120), 5: (1e-30, 1e30), 7: (1e-30, 1e30), 8: (0, 1e30), 17: (-1, 1), 18: (-1, 1), 32: (-0.999, 0.999),
           31: (1, 1e6), 21: (-80, 80), 22: (-80, 80), 0: (-1000, 1000), 1: (-1000, 1000), 2: (-1000, 1000)}
    for op, f in ops.items():
        lo, hi = dom.get(op, (-50, 50))
        if op in (5, 7, 8) or op == 31:
            x = np.exp(rng.uniform(math.log(max(lo, 1e-30)), math.log(hi), 400)).astype(np.float32)
            if op == 31:
                x = (x + 1).astype(np.float32)
        else:
            x = rng.uniform(lo, hi, 400).astype(np.float32)
        got = np.array([L.orc_builtin(op, float(v), 0.0) for v in x], np.float32)
        with np.errstate(all="ignore"):
            want64 = f(x.astype(np.float64))
        want = want64.astype(np.float32)
        ulp = np.abs(np.spacing(want))
        err = np.abs(got.astype(np.float64) - want64)
        assert (err <= 1.0 * ulp + 1e-45).all(), (op, x[np.argmax(err / ulp)], got[np.argmax(err / ulp)], want[np.argmax(err / ulp)])
    # pow and atan2 (two arguments)
    x = np.exp(rng.uniform(-5, 5, 400)).astype(np.float32)
    y = rng.uniform(-8, 8, 400).astype(np.float32)
    got = np.array([L.orc_builtin(3, float(a), float(b)) for a, b in zip(x, y)], np.float32)
    want64 = x.astype(np.float64) ** y.astype(np.float64)
    assert (np.abs(got - want64) <= np.abs(np.spacing(want64.astype(np.float32)))).all()
    got = np.array([L.orc_builtin(16, float(a), float(b)) for a, b in zip(y, x - 3)], np.float32)
    want64 = np.arctan2(y.astype(np.float64), (x - 3).astype(np.float64))
    assert (np.abs(got - want64) <= np.abs(np.spacing(want64.astype(np.float32)))).all()
    # pinned special cases (rm_math.h): pow(negative, 2) is a square (schlick), pow(x,0)=1, pow(0,y>0)=0
    assert L.orc_builtin(3, -0.98, 2.0) == pytest.approx(0.9604, rel=1e-6)
    assert L.orc_builtin(3, -2.0, 3.0) == -8.0
    assert L.orc_builtin(3, 5.0, 0.0) == 1.0 and L.orc_builtin(3, 0.0, 2.0) == 0.0
    assert math.isnan(L.orc_builtin(3, -2.0, 0.5))
    assert L.orc_builtin(3, float(np.float32(0.33333333333)), -1.0) == 3.0


def test_camera_ray_and_march_of_single_pixels():
    """Given the oracle's own jitter values (the RNG is chaotic and cannot be re-derived in another
    precision), the rest of main() - projection, normalisation, the march, depth and colour - is
    recomputed here in float64 and must agree to fp32 accuracy."""
    W, H = 160, 90
    src = scene_source("guide")
    s = rm.default_schema(src, rm.default_custom_settings(src), width=W, height=H)
    s.dof.amount = 0.0                      # no lens offset -> the ray depends on the pixel jitter only
    U = pyoracle.uniforms_from_schema(s, (0.5, 1 / 3))
    checked_hit = checked_sky = 0
    for (px, py) in [(80, 45), (75, 40), (10, 80), (150, 5), (82, 47), (60, 45), (100, 50), (30, 20)]:
        tr = np.zeros(17, np.float32)
        assert L.orc_trace_pixel(b"guide", None, 0, C.byref(U), W, H, px, py, tr.ctypes.data_as(C.c_void_p)) == 0
        o, d, dz, jit, end, steps_taken, depth, rgb = tr[0:3], tr[3:6], tr[6], tr[7:9], tr[9:12], tr[12], tr[13], tr[14:17]
        tc = np.array([(px + 0.5) / W, (py + 0.5) / H]) + jit.astype(np.float64)
        ppp = (tc * 2 - 1) * np.array([W / H, 1.0]) * math.tan(1.5 / 2)
        dn = np.array([ppp[0] + jit[0], ppp[1] + jit[1], 1.0]) * 1.5
        np.testing.assert_allclose(d, dn / np.linalg.norm(dn), atol=3e-7)
        assert dz == pytest.approx(1 / math.sqrt(ppp @ ppp + 1), rel=1e-6)
        np.testing.assert_array_equal(o, np.zeros(3, np.float32))
        # march in float64 from the oracle's ray
        p, dd, st, dep = o.astype(np.float64), d.astype(np.float64), 0.0, 0.0
        for i in range(128):
            sd = sdf_guide(p)
            if sd < 1e11:
                p = p + dd * sd
                dep += float(dz) * sd
            if sd > 1e-4:
                st = float(i)
        if np.linalg.norm(p) > 36:          # sky: position is astronomically far, direction is what matters
            checked_sky += 1
            dy = max(p[1] / np.linalg.norm(p), 0.2)
            np.testing.assert_allclose(rgb, 0.5 * 2 * np.array([0.7, 0.8, 1.0]) * dy, rtol=1e-5)
        else:                               # hit: the surface point agrees to ~1e-4 (fractal surface, fp32 march)
            checked_hit += 1
            assert np.linalg.norm(p - end) < 5e-3
            assert depth == pytest.approx(dep, rel=2e-3)
            assert abs(st - steps_taken) <= 24   # stepsTaken is the chaotic part of the preview shade
    assert checked_sky >= 2 and checked_hit >= 2
