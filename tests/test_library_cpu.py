"""The C-ABI library loads without a GPU and exports every symbol include/rmb.h declares."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import raymarching_engine_b200 as rm
from raymarching_engine_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "rmb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rmb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, set(syms) ^ set(_lib.EXPORTED_SYMBOLS)


def test_no_link_time_dependency_on_the_driver_or_oracle():
    out = subprocess.run(["readelf", "-d", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any("libcuda.so" in n for n in needed), needed      # driver entry points are resolved at run time
    assert not any("oracle" in n for n in needed), needed
    assert any("nvrtc" in n for n in needed), needed


def test_product_sources_never_reference_the_oracle():
    for p in list((ROOT / "raymarching_engine_b200").rglob("*.py")) + list((ROOT / "raymarching_engine_b200" / "csrc").rglob("*.c*")) + \
            list((ROOT / "raymarching_engine_b200" / "csrc").rglob("*.h")):
        text = p.read_text(errors="ignore")
        assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text.replace("the CPU oracle", ""), p


def test_fails_loudly_without_a_gpu_or_with_bad_arguments():
    import torch
    if not torch.cuda.is_available():
        assert rm.load_render_job_context(device=0) is None
        assert "no CPU fallback" in rm.context_error()
        assert rm.load_render_job_group([0, 0]) is None and "no CPU fallback" in rm.group_error()    # device groups likewise
    assert _lib.lib.rmb_ctx_create(0, 3, 2, 16) is None           # rank >= n_ranks
    assert _lib.lib.rmb_render_sample(None, None, None, 0, 0, 1, 1) == _lib.RMB_ERR_INVALID
    assert _lib.lib.rmb_abi_version() == 1
    assert _lib.lib.rmb_group_create(None, 0, 16) is None and _lib.lib.rmb_group_size(None) == 0
    assert _lib.lib.rmb_group_render_sample(None, None, None, 0, 0, 1, 1) == _lib.RMB_ERR_INVALID
    assert _lib.lib.rmb_uniforms_set_frame(None, None) == _lib.RMB_ERR_INVALID
    assert _lib.lib.rmb_stream_write_u32(None, None, 0) == _lib.RMB_ERR_INVALID and _lib.lib.rmb_program_is_live(None) == 0
