#!/usr/bin/env python3
"""Host cost per frame of the public API loop: the default workload at a frame size whose GPU work is negligible
(64x36), so that the wall clock per frame is the Python + ctypes + launch overhead of one do_render_job.
  python tools/host_overhead.py [--frames 400]"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--height", type=int, default=36)
    a = ap.parse_args()
    import torch
    import bench
    import raymarching_engine_b200 as rm
    args = argparse.Namespace(flavour="exact", pipeline="wavefront", scene="guide", step_counts="", spp=1, mode="preview", gather="fused")
    for nctx in (1, 2):
        rig = bench.Rig(rm, torch, args, a.width, a.height, False, None, 0, 1, 0, nctx)
        for i in range(8):
            rig.device_frame(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.frames):
            rig.device_frame(8 + i)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"contexts {nctx}: device_frame enqueue {1e6 * (t1 - t0) / a.frames:7.1f} us/frame, +sync {1e6 * (t2 - t0) / a.frames:7.1f} us/frame")
        for rb in (False, True):
            s, _ = rig.measure_e2e(8, a.frames, False, readback=rb)
            print(f"contexts {nctx}: public API loop readback={rb}: {1e6 * s / a.frames:7.1f} us/frame")
        rig.close()


if __name__ == "__main__":
    main()
