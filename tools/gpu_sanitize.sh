#!/bin/bash
# compute-sanitizer (memcheck) over one small preview + full render through the public API
O=gpurun_out; mkdir -p $O
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, ".")
import raymarching_engine_b200 as rm
ctx = rm.load_render_job_context(device=0)
src = open("scenes/guide.glsl").read(); custom = rm.default_custom_settings(src)
for mode in ("preview", "full"):
    s = rm.default_schema(src, custom, width=77, height=45, renderMode=mode, samplesPerPixel=2, frameid=3 if mode == "preview" else 4)
    if mode == "full": s.lights = [rm.default_light()]
    r = rm.run_job(s, ctx); assert r["success"], r["why"]
ctx.close(); print("sanitize run ok")
PY
compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
