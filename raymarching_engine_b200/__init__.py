"""raymarching_engine_b200 -- B200-native drop-in for the per-pixel raymarch / path-trace hot path of
radian628/raymarching-engine (the WebGL2 fragment shader and its accumulate + display passes).

Public surface mirrors the reference's TypeScript renderer (client/src/renderer/*.tsx):
RenderJobSchema, do_render_job, load_render_job_context, make_presenter, UniformData / u,
get_custom_shader_params.  All rendering goes through libraymarch_b200.so (include/rmb.h); the
import fails if the library has not been built and there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from ._lib import FLAVOUR_EXACT, FLAVOUR_EXACT_ALT, FLAVOUR_FAST
from .executor import (FramebufferInfo, Program, RenderJobContext, ShaderError, builtin_uniforms, context_error,
                       do_render_job, load_render_job_context, make_presenter, render_frames, reset_halton,
                       reset_specialization_history, run_job,
                       upload_sample_uniforms)
from .group import RenderJobGroup, group_error, load_render_job_group
from .halton import halton
from .params import CustomShaderParam, CustomShaderParamError, default_custom_settings, get_custom_shader_params
from .schema import (Camera, Dof, Orthographic, Panoramic, Perspective, PointLight, Render, RenderJobSchema, SunLight,
                     default_light, default_schema)
from .uniforms import UniformData, set_uniform_array, set_uniform_matrix4, set_uniforms, u

__all__ = [n for n in dir() if not n.startswith("_")]
