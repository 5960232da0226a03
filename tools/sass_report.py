#!/usr/bin/env python3
"""SASS evidence without a GPU: lowers + NVRTC-compiles a scene for sm_100a (rmb_compile_only), runs
cuobjdump -sass on the cubin and reports, per kernel, the static instruction mix - and for the march
kernels the mix of the SDF body (the longest straight-line run of the loop), i.e. the per-SDF-step counts of
FFMA / FFMA2 / FADD(2) / FMUL(2) / FRND / MUFU that DESIGN.md argues from.

  python tools/sass_report.py [--scene guide] [--flavour exact|fast] [--kernel rm_wf_march_preview_kernel]
                              [--dump out.sass] [--env RMB_X=1 ...]
"""
from __future__ import annotations

import argparse
import collections
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def compile_scene(scene: str, flavour: str) -> bytes:
    import raymarching_engine_b200 as rm
    from raymarching_engine_b200 import _lib
    p = ROOT / "scenes" / f"{scene}.glsl"
    if not p.exists():
        p = ROOT / "tests" / "fixtures" / "scenes" / f"{scene}.glsl"
    src = p.read_text()
    spec, ns = _lib.make_spec_array(rm.default_custom_settings(src))
    log = C.create_string_buffer(1 << 16)
    cap = 64 << 20
    cubin = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    b = src.encode()
    fl = {"exact": _lib.FLAVOUR_EXACT, "fast": _lib.FLAVOUR_FAST}[flavour]
    st = _lib.lib.rmb_compile_only(b, len(b), fl, spec, ns, log, len(log), cubin, cap, C.byref(n), None, 0)
    if st != 0 or n.value == 0:
        raise SystemExit("compile failed: " + log.value.decode())
    return cubin.raw[:n.value]


def sass_of(cubin: bytes) -> dict:
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
        path = f.name
    try:
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+[^\n]*)", res):
        usage[m.group(1)] = m.group(2)
    return kernels, usage


def opcode(ins: str) -> str:
    t = ins.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


PIPE = {  # coarse pipe classes for the per-step summary
    "fp32": {"FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "FSEL", "FSET", "FSETP", "FCHK", "FSWZADD"},
    "xu": {"MUFU", "FRND", "F2I", "I2F", "F2F", "F2FP", "I2FP"},
    "alu": {"FMNMX", "FMNMX3", "LOP3", "IADD3", "ISETP", "SHF", "PRMT", "SEL", "MOV", "VIADDMNMX", "VIADD", "IMNMX", "PLOP3", "LEA", "IABS", "POPC", "FLO", "BREV", "VIMNMX", "VIMNMX3"},
}


def body_of(ins):
    """the longest run of instructions without control flow (the unrolled SDF body of a march kernel)"""
    ctl = {"BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BREAK", "WARPSYNC", "YIELD", "BMOV", "BAR", "NANOSLEEP"}
    best, cur = [], []
    for _a, i in ins:
        if opcode(i) in ctl:
            if len(cur) > len(best):
                best = cur
            cur = []
        else:
            cur.append(i)
    return best if len(best) >= len(cur) else cur


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="guide")
    ap.add_argument("--flavour", default="exact", choices=["exact", "fast"])
    ap.add_argument("--kernel", default="", help="only this kernel (substring)")
    ap.add_argument("--dump", default="", help="write the selected kernels' SASS here")
    ap.add_argument("--env", nargs="*", default=[])
    a = ap.parse_args()
    for kv in a.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    kernels, usage = sass_of(compile_scene(a.scene, a.flavour))
    dump = []
    print(f"# scene {a.scene}, flavour {a.flavour}, nvcc/NVRTC 12.9 -arch=sm_100a; static SASS counts (cuobjdump -sass)")
    for name, ins in sorted(kernels.items()):
        if a.kernel and a.kernel not in name:
            continue
        c = collections.Counter(opcode(i) for _a, i in ins)
        print(f"\n## {name}: {len(ins)} instructions; {usage.get(name, '')}")
        print("   " + "  ".join(f"{k}:{v}" for k, v in c.most_common(24)))
        if "march" in name:
            b = body_of(ins)
            cb = collections.Counter(opcode(i) for i in b)
            fp = sum(v for k, v in cb.items() if k in PIPE["fp32"])
            fp_lane = fp + sum(v for k, v in cb.items() if k in ("FADD2", "FMUL2", "FFMA2"))
            xu = sum(v for k, v in cb.items() if k in PIPE["xu"])
            print(f"   SDF body (longest straight-line run): {len(b)} instructions, FP32-pipe issue slots {fp} (lane-ops {fp_lane}), XU {xu}")
            print("   body: " + "  ".join(f"{k}:{v}" for k, v in cb.most_common(20)))
        dump.append(f"// ---- {name}\n" + "\n".join(f"/*{ad:04x}*/ {i}" for ad, i in ins))
    if a.dump:
        Path(a.dump).write_text("\n\n".join(dump) + "\n")


if __name__ == "__main__":
    main()
