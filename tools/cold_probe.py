"""Developer probe: C3 frame time with the L2 flushed before every frame (as bench.py does), whole frame and one rank's share of 8."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import torch
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = []
for world in (1, 8):
    ctx.set_shard(0, world)
    for cold in (False, True):
        ms = []
        for _ in range(12):
            if cold:
                flush.zero_()
                torch.cuda.synchronize()
            ctx.render_device(cam, prm)
            ms.append(ctx.sync().gpu_ms)
        ms = sorted(ms[2:])
        out.append(f"1/{world} {'cold' if cold else 'warm'} L2: median {ms[len(ms) // 2]:.3f} min {ms[0]:.3f} ms")
print(" | ".join(out), flush=True)
