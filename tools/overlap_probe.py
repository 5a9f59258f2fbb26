"""Developer probe: C3 frame time with and without the side-stream overlap, whole frame and one rank's share of N."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
for world in (1, 2, 4, 8):
    ctx.set_shard(0, world)
    out = []
    for ov in (False, True):
        ctx.set_overlap(ov)
        ms = []
        for _ in range(8):
            ctx.render_device(cam, prm)
            ms.append(ctx.sync().gpu_ms)
        out.append(min(ms[2:]))
    print(f"world {world}: rank-0 frame {out[0]:.3f} ms sequential, {out[1]:.3f} ms overlapped  (ideal {2.8 / world:.3f})", flush=True)
