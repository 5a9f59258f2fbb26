"""Developer probe: device-resident C3 frame time for lanes x batches shapes (RTB200_GRID_MULT selects the traversal grid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200
from rtb200 import standin
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
out = []
for shape in [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]:
    ctx.set_pipeline(*shape)
    ms = []
    for _ in range(8):
        ctx.render_device(cam, prm)
        ms.append(ctx.sync().gpu_ms)
    out.append(f"{shape[0]}x{shape[1]}: {min(ms[2:]):.3f}")
print(" | ".join(out), flush=True)
