"""Developer probe: rt_render into page-locked memory for the BASELINE configs C1-C4 (wall ms per frame, best of 8)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import rtb200
from rtb200 import standin
from util import Golden
ctx = rtb200.Context(0)
cam = rtb200.make_camera()
cases = [("C1 cornell 1024^2 d3", Golden("cornell_c1_256").scene, rtb200.make_params(1024, 1024, 3)),
         ("C2 teapot 1920x1080 d0", Golden("teapot_c2_256x144").scene, rtb200.make_params(1920, 1080, 0)),
         ("C3 dragon* 3840x2160 d3", standin.dragon_standin_scene(), rtb200.make_params(3840, 2160, 3)),
         ("C4 cornell 2048^2 sph64 d5", Golden("cornell_c4_96").scene, rtb200.make_params(2048, 2048, 5, sphere_rays=64))]
out = []
for name, sc, prm in cases:
    ctx.upload_scene(sc, rtb200.BVH_SAH_HOST)
    pinned = torch.empty(prm.width * prm.height * 3, dtype=torch.float32).pin_memory()
    wall = []
    for _ in range(8):
        t0 = time.perf_counter()
        st = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
        wall.append(1e3 * (time.perf_counter() - t0))
    ctx.render_device(cam, prm)
    dev = ctx.sync().gpu_ms
    out.append(f"{name}: host {min(wall[2:]):.3f} ms (device-resident {dev:.3f} ms, {st.kernel_launches} launches)")
print("\n".join(out), flush=True)
