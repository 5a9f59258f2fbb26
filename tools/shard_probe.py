"""Developer probe: C3 frame time of ONE rank's share when the frame is split over 1 / 2 / 4 / 8 ranks (run on one GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
out = []
for world in (1, 2, 4, 8):
    ctx.set_shard(0, world)
    ms = []
    for _ in range(10):
        ctx.render_device(cam, prm)
        ms.append(ctx.sync().gpu_ms)
    out.append(f"1/{world}: {min(ms[2:]):.3f} ms")
print(" | ".join(out), flush=True)
