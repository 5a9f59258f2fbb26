"""Developer probe: C3 device-resident frame time for pipeline shapes (lanes x batches per frame) and grid multipliers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx.set_shard(0, world)
out = []
for lanes, batches in ((0, 1), (1, 1), (2, 2), (2, 4), (3, 3), (4, 4), (3, 6), (4, 8)):
    ctx.set_pipeline(lanes, batches, 1 << 14)
    ms = []
    for _ in range(12):
        ctx.render_device(cam, prm)
        ms.append(ctx.sync().gpu_ms)
    out.append(f"{lanes}x{batches} {np.median(ms[2:]):.3f}")
print(f"world {world} grid_mult {os.environ.get('RTB200_GRID_MULT', '-')}: " + " | ".join(out), flush=True)
