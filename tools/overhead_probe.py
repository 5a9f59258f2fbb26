"""Developer probe: fixed cost of a frame — C3's scene with the camera turned away (every ray misses the root box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
for name, cam in (("looking away", rtb200.make_camera(look_at=(50.0, 0.0, 0.0), dist=3.0)), ("normal", rtb200.make_camera())):
    for w, h in ((3840, 2160), (1024, 1024), (256, 256)):
        for depth in (0, 3):
            prm = rtb200.make_params(w, h, depth)
            for lanes, overlap in ((0, True), (1, True), (1, False)):
                ctx.set_pipeline(lanes, 1 if lanes else 0)
                ctx.set_overlap(overlap)
                ms = []
                for _ in range(12):
                    ctx.render_device(cam, prm)
                    st = ctx.sync()
                    ms.append(st.gpu_ms)
                print(f"{name:13s} {w}x{h} depth {depth} lanes {lanes} overlap {int(overlap)}: {min(ms[2:]):.3f} ms, {st.kernel_launches} launches, {st.batches} batches", flush=True)
