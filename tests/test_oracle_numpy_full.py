"""Independent cross-check of the oracle's FULL (path-tracing) branch: a float64 restatement of
raymarcher.frag:178-205 (camera + thin lens) and :246-387 (bounce loop, lights, blend), written here
from the reference GLSL with plain Python/numpy scalars - no vec types, no glsl_rt.h / rm_math.h, no
fused operations - and replayed against the event log of the C++ oracle (orc_trace_full_pixel).

What is taken from the oracle's log as INPUT (chaotic, not reproducible in another precision; each is
covered by its own test): the RNG draws (gold_noise amplifies 1 ulp of tan), the end
points of castRay (the fractal surface; the march loop itself is re-derived in test_oracle_numpy.py),
the finite-difference normal (a difference of fp32 SDF values 1e-5 apart) and the subsurface probe's
SDF value.  What is RE-DERIVED and compared: the order in which the shader consumes its random
numbers, Box-Muller and the sphere sample, the lens and the camera ray, the fog free path, emission,
the subsurface sample, the branch taken (:278 / :284 / :300), the diffuse flip, reflect + Rodrigues +
Schlick, the :334 offset, the bounce-0 attachments, the shadow test and the light term, and the
final blend - i.e. everything of the path tracer that the preview-branch cross-check does not reach.
"""
import ctypes as C
import math

import numpy as np
import pytest

import pyoracle
import raymarching_engine_b200 as rm
from conftest import scene_source

L = pyoracle.lib()
f32 = lambda x: float(np.float32(x))   # noqa: E731  (a GLSL literal as the shader sees it)

PAYLOAD = {1: 1, 2: 2, 3: 6, 10: 7, 11: 3, 12: 3, 13: 4, 14: 12, 15: 8, 16: 4, 17: 3, 18: 4}


def trace(scene, schema, noise, px, py):
    U = pyoracle.uniforms_from_schema(schema, noise)
    cu = pyoracle.flatten_custom(scene, schema.customShaderParameters)
    buf = np.zeros(1 << 14, np.float32)
    L.orc_trace_full_pixel.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = L.orc_trace_full_pixel(scene.encode(), cu.ctypes.data_as(C.c_void_p), int(cu.size), C.byref(U), schema.render.width,
                               schema.render.height, px, py, buf.ctypes.data_as(C.c_void_p), len(buf))
    assert 0 < n <= len(buf)
    ev, i = [], 0
    while i < n:
        tag = int(buf[i])
        k = PAYLOAD[tag]
        ev.append((tag, buf[i + 1:i + 1 + k].astype(np.float64)))
        i += 1 + k
    return ev


class Log:
    def __init__(self, ev):
        self.ev, self.i = ev, 0

    def take(self, tag):
        t, v = self.ev[self.i]
        assert t == tag, f"event {self.i}: the shader's next event is {t}, the restatement expects {tag}"
        self.i += 1
        return v

    def peek(self):
        return self.ev[self.i][0] if self.i < len(self.ev) else None


def gmax(a, b):
    """GLSL max with the pinned NaN behaviour (a NaN operand is dropped, IEEE maxNum)"""
    return b if math.isnan(a) else a if math.isnan(b) else max(a, b)


def gmin(a, b):
    return b if math.isnan(a) else a if math.isnan(b) else min(a, b)


def gsign(x):
    """GLSL sign as the comparison chain every implementation lowers it to: sign(NaN) = 0"""
    return 1.0 if x > 0 else -1.0 if x < 0 else 0.0


def norm(v):
    """length(): unscaled sqrt(x*x + y*y + z*z) (the pin of SURVEY.md H2) - the sum of squares overflows
    where fp32 does (escaped rays reach 1e19), which decides e.g. the sky colour of guide.glsl:85"""
    d = float(v @ v)
    return math.inf if d > 3.4028234663852886e38 else math.sqrt(d)


def normalize(v):
    with np.errstate(all="ignore"):
        return v / norm(v)


def close(got, want, what, rtol=5e-5, atol=2e-6):
    got, want = np.asarray(got, float), np.asarray(want, float)
    scale = max(1.0, float(np.max(np.abs(want[np.isfinite(want)]))) if np.isfinite(want).any() else 1.0)
    assert np.allclose(got, want, rtol=rtol, atol=atol * scale, equal_nan=True), (what, got, want)


# ---- guide.glsl material functions (client/public/examples/guide.glsl:51-88), float64 ----
def guide_materials(p, fractal_color):
    far35, far36 = norm(p) > 35.0, norm(p) > 36.0
    diffuse = np.zeros(3) if far35 else np.asarray(fractal_color, float)
    specular = np.zeros(3) if far35 else np.full(3, f32(0.6))
    d = gmax(float(normalize(p)[1]), f32(0.2))
    emission = np.array([f32(0.7), f32(0.8), 1.0]) * d * 2.0 if far36 else np.zeros(3)
    return dict(diffuse=diffuse, specular=specular, roughness=f32(0.2), sss=11111115.0, sss_color=np.ones(3), ior=100.0, emission=emission)


def replay(ev, schema, noise):
    """raymarcher.frag main(), full branch, in float64 on the oracle's random numbers"""
    log = Log(ev)
    W, H = schema.render.width, schema.render.height
    PI = f32(3.141592)

    def uniform_sample():
        return float(log.take(1)[0])

    def box_muller():                                  # :80-89
        u1, u2 = log.take(2)
        with np.errstate(all="ignore"):
            r = math.sqrt(-2.0 * math.log(u1)) if u1 > 0 else math.inf
        return np.array([r * math.cos(2.0 * PI * u2), r * math.sin(2.0 * PI * u2)])

    def sphere_sample():                               # :96-101, arguments left to right
        a = box_muller()
        b = box_muller()[0]
        return normalize(np.array([a[0], a[1], b]))

    fractal_color = schema.customShaderParameters["fractalColor"].data
    position = np.asarray(schema.camera.position, float)
    rot = np.asarray(schema.camera.rotation, float).reshape(4, 4).T      # column-major upload
    exposure = f32(schema.render.exposure / schema.render.samplesPerPixel)

    # ---- camera, :180-193 (perspective)
    jit = np.array([uniform_sample(), uniform_sample()]) / np.array([W, H], float)
    texcoord = np.array([(schema._px + 0.5) / W, (schema._py + 0.5) / H])
    tc2 = texcoord + jit
    dof_offset = sphere_sample() * f32(schema.dof.amount)
    ray_p = position + dof_offset
    ppp = (tc2 * 2.0 - 1.0) * np.array([f32(W / H), 1.0]) * math.tan(f32(schema.camera.mode.fov) / 2.0)
    not_normalized = (rot @ np.array([ppp[0] + jit[0], ppp[1] + jit[1], 1.0, 0.0]))[:3]
    ray_d = normalize(not_normalized * f32(schema.dof.distance) - dof_offset)
    cam = log.take(3)
    close(ray_p, cam[:3], "camera origin")
    close(ray_d, cam[3:], "camera direction", atol=2e-6)

    albedo, light = np.ones(3), np.zeros(3)
    branches = []
    counts = schema.reflectionIterationCounts
    for i in range(len(counts)):
        start = log.take(10)
        assert int(start[0]) == i
        close(ray_p, start[1:4], f"bounce {i} start position")
        close(ray_d, start[4:7], f"bounce {i} start direction", atol=5e-6)
        old_p = ray_p
        ray_p = log.take(11)                                             # castRay end point (input)
        x = uniform_sample()
        with np.errstate(all="ignore"):
            path_length = np.float64(-math.log(1.0 - x)) / np.float64(f32(schema.fogDensity))     # invExpDist :148-150
        m = guide_materials(ray_p, fractal_color)
        light = light + albedo * m["emission"]
        normal = log.take(12)                                            # finite-difference normal (input)
        sub_sample = -1.0 / m["sss"] * math.log(1.0 - uniform_sample())
        sub_dir = normalize(normalize(sphere_sample()))                  # mix(a, b, 1.0) = b
        with np.errstate(all="ignore"):
            sub_dir = sub_dir * -gsign(float(sub_dir @ normal))
        sub_pos = ray_p + sub_dir * sub_sample
        prev_albedo, diffuse, specular, prev_d = albedo, m["diffuse"], m["specular"], ray_d
        with np.errstate(all="ignore"):
            escaped = (norm(old_p - ray_p) > path_length) or bool(np.isinf(ray_p).any()) or bool(np.isnan(ray_p).any())
        if escaped:                                                       # :278-283
            branches.append("fog/escape")
            ray_p = old_p + gmin(float(path_length), 1000000.0) * ray_d
            ray_d = sphere_sample()
            diffuse, specular, prev_d = np.ones(3), np.ones(3), ray_d
        else:
            probe = log.take(13)
            close(sub_pos, probe[1:4], "subsurface probe position", atol=3e-6)
            if probe[0] > f32(0.001):                                    # :284-288
                branches.append("subsurface")
                albedo = albedo * m["sss_color"]
                ray_p = sub_pos
                ray_d = normalize(sphere_sample())
            else:
                db, sb = norm(diffuse), norm(specular)
                prob = (1.0 - sb / db / 2.0) if db > sb else (db / sb / 2.0)
                if uniform_sample() < prob:                              # :300-319
                    branches.append("diffuse")
                    albedo = albedo * diffuse
                    new_d = sphere_sample()
                    ray_d = gsign(float(normal @ new_d)) * new_d
                else:                                                    # :322-330
                    branches.append("specular")
                    cos_t = -float(ray_d @ normal)
                    r0 = ((1.0 - m["ior"]) / (1.0 + m["ior"])) ** 2
                    fres = r0 + (1.0 - r0) * (1.0 - cos_t) ** 5
                    albedo = albedo * specular * gmin(gmax(fres, 0.0), 1.0)
                    rand_vec = sphere_sample()
                    ray_d = ray_d - 2.0 * float(normal @ ray_d) * normal          # reflect
                    axis = normalize(np.cross(rand_vec, ray_d))
                    theta = m["roughness"] * uniform_sample()
                    c = math.cos(theta)
                    sn = math.sqrt(max(1.0 - c * c, 0.0))
                    ray_d = ray_d * c + np.cross(axis, ray_d) * sn + axis * float(axis @ ray_d) * (1.0 - c)   # rodrigues :61-65
        ray_p = ray_p + ray_d * f32(0.001)                               # :334
        after = log.take(14)
        close(ray_p, after[0:3], f"bounce {i} position after :334", atol=3e-6)
        close(ray_d, after[3:6], f"bounce {i} direction", atol=2e-5)
        close(albedo, after[6:9], f"bounce {i} albedo")
        close(light, after[9:12], f"bounce {i} light before the light loop")
        if i == 0:                                                       # :336-352
            depth = gmin(gmax(norm(ray_p - position), f32(0.00001)), 100000000.0)
            normal = np.where(np.isfinite(normal), normal, 0.0)      # :339-341 scrub the variable the light loop reads
            n_out = normal
            with np.errstate(all="ignore"):
                dof_radius = gmin(gmax(f32(schema.dof.amount) * abs(depth - f32(schema.dof.distance)) / depth, 0.0), 1.0)
            aux = log.take(15)
            close(n_out, aux[0:3], "normal attachment")
            close(dof_radius, aux[3], "dofRadius", rtol=2e-4)
            close(albedo, aux[4:7], "albedo attachment")
            close(depth, aux[7], "depth attachment", rtol=1e-5)
        for j, lt in enumerate(schema.lights):                           # :354-373
            lp = np.asarray(lt.position, float)
            lc = np.array([f32(c) for c in lt.color])
            adj = lp + sphere_sample() * f32(lt.size)
            to_light = normalize(adj - ray_p)
            res = log.take(16)
            assert int(res[0]) == j
            result = res[1:4]
            with np.errstate(all="ignore"):
                d_res, d_pos = norm(result - adj), norm(ray_p - adj)
                lit = d_res >= d_pos
            if np.isfinite(d_res) and abs(d_res - d_pos) < 1e-4 * max(1.0, d_pos):
                return None                                              # borderline visibility: fp32 decides, skip this pixel
            if lit:
                refl = prev_d - 2.0 * float(normal @ prev_d) * normal
                r = gmax(0.0, float(to_light @ refl))
                rough = m["roughness"]
                light = light + prev_albedo * diffuse * lc * gmax(0.0, float(to_light @ normal)) \
                    + prev_albedo * specular * lc * rough * rough / (f32(3.14159265) * (r * r * (rough * rough - 1.0) + 1.0) ** 2)
            close(light, log.take(17), f"bounce {i} light {j}", rtol=2e-4)
    frag = log.take(18)                                                  # :379-387 on zero previous texels
    if schema.render.blendMode == "additive":
        want = np.array([*(light * exposure), 1.0])
    else:
        f = f32(schema.render.blendWithPreviousFrameFactor)
        want = np.array([*(light * exposure), 1.0]) * (1.0 - f)
    close(want, frag, "fragColor", rtol=3e-4)
    assert log.peek() is None
    return branches


@pytest.mark.parametrize("fog,blend", [(0.0, "additive"), (0.05, "mix")])
def test_full_branch_replayed_in_float64(fog, blend):
    W, H = 96, 54
    src = scene_source("guide")
    s = rm.default_schema(src, rm.default_custom_settings(src), width=W, height=H, renderMode="full", blendMode=blend)
    s.lights = [rm.default_light(), rm.default_light()]
    s.lights[1].position = (2.0, 3.0, 4.0)
    s.lights[1].size = 0.5
    s.lights[1].color = (0.5, 1.0, 2.0)
    s.fogDensity = fog
    seen, done = {}, 0
    for py in range(1, H, 2):
        for px in range(1, W, 3):
            s._px, s._py = px, py
            noise = (0.5, 1.0 / 3.0)
            br = replay(trace("guide", s, noise, px, py), s, noise)
            if br is None:
                continue
            done += 1
            for b in br:
                seen[b] = seen.get(b, 0) + 1
    assert done >= 700
    # every branch of the bounce logic was exercised (fog 0: no subsurface hits in this scene -
    # sceneSubsurfaceScattering is 1.1e7, the probe lands ~1e-7 from the surface - so "subsurface" is optional)
    assert seen.get("fog/escape", 0) > 20 and seen.get("diffuse", 0) > 20 and seen.get("specular", 0) > 5, seen
