"""Developer probe: wall time of rt_render (pinned host buffer) against the device time between the frame's first and last event."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import torch
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
sc = standin.dragon_standin_scene()
ctx.upload_scene(sc, rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
pinned = torch.empty(3840 * 2160 * 3, dtype=torch.float32).pin_memory()
for shape in [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(0, 0)]:
    ctx.set_pipeline(*shape)
    wall, dev, setup = [], [], []
    for _ in range(12):
        t0 = time.perf_counter()
        ctx.set_materials(sc.mats)
        ctx.set_lights(sc.point_lights, sc.sphere_lights)
        t1 = time.perf_counter()
        st = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
        t2 = time.perf_counter()
        wall.append(1e3 * (t2 - t1)); dev.append(st.gpu_ms); setup.append(1e3 * (t1 - t0))
    k = min(range(2, 12), key=lambda i: wall[i])
    print(f"lanes x batches {shape}: rt_render wall {wall[k]:.3f} ms, device first-to-last event {dev[k]:.3f} ms, materials + lights upload {setup[k]:.3f} ms, {st.kernel_launches} launches", flush=True)
