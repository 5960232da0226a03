"""The render-job description.  Mirrors /root/reference/client/src/renderer/RenderJobSchema.tsx:1-86
field for field (camelCase names kept so a job written for the reference reads the same here)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Union

from .uniform_types import UniformData

Vec3 = Sequence[float]


@dataclass
class PointLight:                       # RenderJobSchema.tsx:5-10
    position: Vec3
    color: Vec3
    size: float
    type: str = "point"


@dataclass
class SunLight:                         # RenderJobSchema.tsx:11-15 (treated as a point at `direction`, size 0:
    direction: Vec3                     #  RenderJobExecutor.tsx:280, :289)
    color: Vec3
    type: str = "sun"


@dataclass
class Perspective:                      # RenderJobSchema.tsx:50-53
    fov: float
    type: str = "perspective"


@dataclass
class Orthographic:                     # RenderJobSchema.tsx:54-57
    size: float
    type: str = "orthographic"


@dataclass
class Panoramic:                        # RenderJobSchema.tsx:58-62
    angleX: float = 6.283185307179586
    angleY: float = 3.141592653589793
    type: str = "panoramic"


IDENTITY4 = (1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0)


@dataclass
class Dof:                              # RenderJobSchema.tsx:38-42
    amount: float = 0.01
    distance: float = 1.5
    showFocusedArea: bool = False


@dataclass
class Camera:                           # RenderJobSchema.tsx:44-63
    position: Vec3 = (0.0, 0.0, 0.0)
    motion: Vec3 = (0.0, 0.0, 0.0)      # never consumed by the reference
    rotation: Sequence[float] = IDENTITY4   # gl-matrix mat4, column-major
    mode: Union[Perspective, Orthographic, Panoramic] = field(default_factory=lambda: Perspective(1.5))


@dataclass
class Render:                           # RenderJobSchema.tsx:65-83
    samplesPerPixel: int = 1
    exposure: float = 0.5
    subdivisions: int = 1
    width: int = 1280
    height: int = 720
    frameid: int = 0
    blendWithPreviousFrameFactor: float = 0.9
    sampleYieldInterval: int = 1
    blendMode: str = "additive"         # "additive" | "mix"
    renderMode: str = "preview"         # "full" | "preview"


@dataclass
class RenderJobSchema:                  # RenderJobSchema.tsx:17-86
    sdfShaderSource: str
    reflectionIterationCounts: List[float] = field(default_factory=lambda: [128, 128, 64, 32, 32])
    normalDelta: float = 0.00001        # never consumed (the shader uses the literal, raymarcher.frag:264)
    customShaderParameters: Dict[str, UniformData] = field(default_factory=dict)
    fogDensity: float = 0.0
    time: float = 0.0                   # never consumed
    timeDelta: float = 0.0              # never consumed
    dof: Dof = field(default_factory=Dof)
    camera: Camera = field(default_factory=Camera)
    render: Render = field(default_factory=Render)
    lights: List[Union[PointLight, SunLight]] = field(default_factory=list)


def default_light() -> PointLight:
    """The light the UI's "Add light" button creates (settings/LightSettings.tsx:82-87) after the
    colour scaling of index.tsx:174: rgb * strength / 256 with rgb = 255, strength = 3."""
    c = 255 * 3 / 256
    return PointLight(position=(0.0, 0.0, 0.0), color=(c, c, c), size=0.0)


def default_schema(scene_source: str, custom=None, **render_kw) -> RenderJobSchema:
    """The job the reference app builds on start-up (index.tsx:121-182 with the settings defaults of
    index.tsx:308-335): preview mode, no lights, 1 spp, exposure 0.5, additive blending."""
    s = RenderJobSchema(sdfShaderSource=scene_source, customShaderParameters=dict(custom or {}))
    for k, v in render_kw.items():
        setattr(s.render, k, v)
    return s
