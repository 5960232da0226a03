"""integration/: the N-API addon a maintainer of the reference would build cannot be compiled against a
real Node.js here (none in the image), but it must at least be valid C against include/rmb.h and the
N-API signatures (integration/stub/node_api.h), and bind every entry point the TypeScript shim calls."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_napi_addon_is_valid_c_against_the_header():
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"),
                        "-I", str(ROOT / "integration" / "stub"), str(ROOT / "integration" / "rmb_napi.c")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_ts_shim_only_calls_exported_addon_functions():
    c = (ROOT / "integration" / "rmb_napi.c").read_text()
    ts = (ROOT / "integration" / "renderer" / "RenderJobExecutorB200.ts").read_text()
    exported = set(re.findall(r'EXPORT\("(\w+)"', c))
    used = set(re.findall(r"\brmb\.(\w+)\b", ts))          # called directly or picked by `gl.group ? rmb.a : rmb.b`
    used -= {"h", "node"}                                        # include/rmb.h in a comment, native/rmb.node in the require()
    assert len(used) >= 20 and used <= exported, used - exported
    # device groups: every single-device call the shim makes has its group twin bound as well
    for name in ("ProgramGet", "UniformSet", "UniformSetArray", "UniformMatrix4", "FbAcquire", "FbRelease", "RenderSample", "Present"):
        assert name[0].lower() + name[1:] in exported and "group" + name in exported, name


def test_header_is_valid_c_and_cpp():
    for lang, std in (("c", "c11"), ("c++", "c++17")):
        p = subprocess.run(["gcc", "-x", lang, f"-std={std}", "-Wall", "-Werror", "-fsyntax-only", str(ROOT / "include" / "rmb.h")], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
