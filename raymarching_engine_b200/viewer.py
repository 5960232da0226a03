"""The interactive camera / frame state machine of the reference's real-time mode, without the browser:
/root/reference/client/src/index.tsx:61-283 (`realtimeMode`).  Pointer-lock mouse look, WASD + shift /
space fly camera, the "new frame" rule that bumps `frameid` (and thereby clears the accumulators) one
loop AFTER the camera moved, and the per-frame sample counter that feeds the presenter's brightness.

Matrices follow gl-matrix 3.4.3 (`mat4.rotate`, `vec3.transformMat4`), which stores Float32Array
column-major; arithmetic is done in float32 like the typed arrays the reference uses."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

F = np.float32
EPSILON = 0.000001          # glMatrix.EPSILON


def mat4_identity() -> np.ndarray:
    return np.eye(4, dtype=F).reshape(16).copy()


def mat4_rotate(a: np.ndarray, rad: float, axis) -> np.ndarray:
    """gl-matrix mat4.rotate(out, a, rad, axis) (mat4.js): a * R(rad, axis); returns a new matrix (the
    reference passes the same array as out and a)."""
    x, y, z = (float(v) for v in axis)
    ln = math.hypot(x, y, z)
    if ln < EPSILON:
        return a.copy()
    x, y, z = x / ln, y / ln, z / ln
    s, c = math.sin(rad), math.cos(rad)
    t = 1 - c
    a = a.astype(F)
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23 = (float(v) for v in a[:12])
    b00, b01, b02 = x * x * t + c, y * x * t + z * s, z * x * t - y * s
    b10, b11, b12 = x * y * t - z * s, y * y * t + c, z * y * t + x * s
    b20, b21, b22 = x * z * t + y * s, y * z * t - x * s, z * z * t + c
    out = a.copy()
    out[0:4] = [a00 * b00 + a10 * b01 + a20 * b02, a01 * b00 + a11 * b01 + a21 * b02, a02 * b00 + a12 * b01 + a22 * b02, a03 * b00 + a13 * b01 + a23 * b02]
    out[4:8] = [a00 * b10 + a10 * b11 + a20 * b12, a01 * b10 + a11 * b11 + a21 * b12, a02 * b10 + a12 * b11 + a22 * b12, a03 * b10 + a13 * b11 + a23 * b12]
    out[8:12] = [a00 * b20 + a10 * b21 + a20 * b22, a01 * b20 + a11 * b21 + a21 * b22, a02 * b20 + a12 * b21 + a22 * b22, a03 * b20 + a13 * b21 + a23 * b22]
    return out.astype(F)


def vec3_transform_mat4(v, m: np.ndarray) -> Tuple[float, float, float]:
    """gl-matrix vec3.transformMat4(out, a, m): m * (a, 1) with perspective divide."""
    x, y, z = (float(c) for c in v)
    m = [float(c) for c in m]
    w = m[3] * x + m[7] * y + m[11] * z + m[15]
    w = w or 1.0
    return (float(F((m[0] * x + m[4] * y + m[8] * z + m[12]) / w)), float(F((m[1] * x + m[5] * y + m[9] * z + m[13]) / w)),
            float(F((m[2] * x + m[6] * y + m[10] * z + m[14]) / w)))


@dataclass
class RealtimeController:
    """State of `realtimeMode`.  Feed it input events, call begin_loop() once per animation frame to get
    (frameid, samples_rendered_so_far) for the job and the presenter, then end_loop() after the job."""
    camera_speed: float = 0.01
    camera_rotation: np.ndarray = field(default_factory=mat4_identity)                 # index.tsx:74
    viewer_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])     # settings().viewerPosition, index.tsx:333
    frameid: int = 0                                                                    # index.tsx:76
    samples_rendered_so_far: int = 0                                                    # index.tsx:118
    pointer_locked: bool = True
    requesting_new_frame: bool = False                                                  # settings().requestingNewFrame
    _mouse_has_moved: int = 0
    _previously_switched: bool = False
    _keys: Dict[str, bool] = field(default_factory=dict)
    _accel: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    _job_frameid: int = 0

    def mouse_move(self, movement_x: float, movement_y: float) -> None:                # index.tsx:78-95
        if not self.pointer_locked:
            return
        self.camera_rotation = mat4_rotate(self.camera_rotation, 0.004 * movement_x, (0.0, 1.0, 0.0))
        self.camera_rotation = mat4_rotate(self.camera_rotation, 0.004 * movement_y, (1.0, 0.0, 0.0))
        self._mouse_has_moved = 5

    def key(self, name: str, down: bool) -> None:                                       # index.tsx:97-105
        if down and not self.pointer_locked:
            return
        self._keys[name.lower()] = down

    def begin_loop(self) -> Tuple[int, int]:                                            # index.tsx:120-233
        """One iteration of loop() up to the doRenderJob call: returns (frameid of THIS loop's job, samplesRenderedSoFar
        handed to makePresenter).  The reference builds testRenderJob - render.frameid included, index.tsx:121-182 -
        at the top of loop(), BEFORE the new-frame rule increments frameid (:221-231): the loop that resets the sample
        counter to 1 still accumulates into the OLD frame's buffers, and the fresh buffers are first drawn one loop later,
        presented with brightness 1/2.  Reproduced as is (bug-compatible; tests/test_viewer_*)."""
        self._job_frameid = self.frameid                     # what testRenderJob captured
        sp = self.camera_speed
        ax = (sp if self._keys.get("d") else 0.0) - (sp if self._keys.get("a") else 0.0)
        ay = (sp if self._keys.get(" ") else 0.0) - (sp if self._keys.get("shift") else 0.0)
        az = (sp if self._keys.get("w") else 0.0) - (sp if self._keys.get("s") else 0.0)
        self._accel = (ax, ay, az)
        should_switch = math.hypot(ax, ay, az) != 0 or self._mouse_has_moved > 0
        if self._previously_switched or self.requesting_new_frame:
            self.requesting_new_frame = False
            self.samples_rendered_so_far = 0
            self.frameid += 1
            self._previously_switched = False
        self.samples_rendered_so_far += 1
        if should_switch:
            self._previously_switched = True
        return self._job_frameid, self.samples_rendered_so_far

    def end_loop(self) -> None:                                                         # index.tsx:267-279
        self._mouse_has_moved -= 1
        accel = vec3_transform_mat4(self._accel, self.camera_rotation)
        if math.hypot(*accel) != 0:
            self.viewer_position = [float(F(p + a)) for p, a in zip(self.viewer_position, accel)]

    def apply_to(self, schema) -> None:
        """what index.tsx:121-182 copies into the job every frame"""
        schema.camera.position = tuple(self.viewer_position)
        schema.camera.rotation = tuple(float(v) for v in self.camera_rotation)
        schema.render.frameid = self._job_frameid
