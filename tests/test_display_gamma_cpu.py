"""The present pass's gamma + RGBA8 quantisation as a table search (csrc/display.cu gammaByte, device_src/
display_gamma_table.inc): the committed thresholds are regenerated from the shared binary64 pow (rm_math.h, host build)
and checked against `floor(clamp(pow(x, 1/2.2), 0, 1) * 255 + 0.5)` for EVERY non-negative float - 2^31 values, a few
seconds on all cores - which also proves the monotonicity the table relies on (tools/gen_gamma_table.cpp)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DEV = ROOT / "raymarching_engine_b200" / "csrc" / "device_src"


def test_gamma_table_is_the_pow_for_every_float(tmp_path):
    exe = tmp_path / "gen_gamma"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-pthread", "-I", str(DEV),
                    str(ROOT / "tools" / "gen_gamma_table.cpp"), "-o", str(exe)], check=True)
    regenerated = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert regenerated == (DEV / "display_gamma_table.inc").read_text()
    r = subprocess.run([str(exe), "--verify", str(DEV / "display_gamma_table.inc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "2139095041 values, 0 mismatches" in r.stdout
