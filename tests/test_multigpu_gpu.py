"""Multi-GPU row-tile sharding on real GPUs (needs >= 2 devices; skipped otherwise): two ranks render
interleaved 16-row tiles of one frame and assemble it on rank 0 with the fused peer-store gather
(display kernel -> rank 0's buffer over CUDA IPC / NVLink); the result must equal the single-GPU frame."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RM_ROOT"])
import raymarching_engine_b200 as rm
from raymarching_engine_b200.sharding import FusedTileGather
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H = 200, 120
src = open(os.path.join(os.environ["RM_ROOT"], "scenes", "guide.glsl")).read()
custom = rm.default_custom_settings(src)
ctx = rm.load_render_job_context(device=rank, rank=rank, n_ranks=world, tile_rows=16)
g = FusedTileGather([ctx], W, H, dist, slots=2, blur=True)
ok = True
for frame, mode in enumerate(["preview", "full"]):
    s = rm.default_schema(src, custom, width=W, height=H, renderMode=mode, frameid=10 + frame)
    if mode == "full":
        s.lights = [rm.default_light()]
        s.dof.amount = 0.05          # visible depth-of-field blur: the display pass reads neighbour tiles
    rm.reset_halton()
    fb = ctx.fbo.create(W, H, s.render.frameid)
    if mode == "preview":
        g.aim(ctx, frame)            # display kernel stores straight into rank 0's frame
        out = rm.run_job(s, ctx)
        assert out["success"], out["why"]
        g.complete(ctx)
    else:
        rm._lib.lib.rmb_ctx_set_gather_target(ctx.handle, None, 0)
        # accumulators of this rank's tiles; no present on the member context (a tile-owning context refuses to present
        # a frame that may blur: the display pass would read rows it does not hold)
        gen = rm.do_render_job(s, ctx)(lambda *a: None)
        try:
            while True:
                next(gen)
        except StopIteration as stop:
            out = stop.value
        assert out["success"], out["why"]
        try:
            ctx.present(fb, 1.0)
            raise SystemExit("present of a blurred tile frame must be refused")
        except RuntimeError as e:
            assert "neighbour rows" in str(e), e
        g.scatter(ctx, fb, frame)    # colour + normal/dofRadius rows -> rank 0's full-frame planes
        g.complete(ctx)
        if rank == 0:
            g.display_assembled(ctx, frame, 1.0)
    ctx.sync()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        got = g.frame_tensor(frame).cpu().numpy()
        one = rm.load_render_job_context(device=0)
        rm.reset_halton()
        s.render.frameid = 100 + frame
        want = rm.run_job(s, one)["rgba8"].copy()
        one.close()
        same = bool(np.array_equal(got, want))
        print(f"{mode}: fused gather identical to single GPU = {same}", flush=True)
        ok = ok and same
g.close()
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_fused_tile_gather_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RM_ROOT=str(ROOT))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "preview: fused gather identical to single GPU = True" in p.stdout
    assert "full: fused gather identical to single GPU = True" in p.stdout
