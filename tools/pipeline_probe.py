"""Developer probe: C3 frame time (device-resident) and rt_render time (pinned host buffer) vs pipeline lanes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import numpy as np
import torch
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
pinned = torch.empty(3840 * 2160 * 3, dtype=torch.float32).pin_memory()
ref = None
for world in (1,):
    ctx.set_shard(0, world)
    for lanes, nb in ((1, 1), (1, 2), (1, 3), (2, 2), (2, 4), (2, 6), (3, 3), (3, 6), (3, 9), (4, 4), (4, 8), (2, 3), (1, 4)):
        minb = 1 << 14
        ctx.set_pipeline(lanes, nb, minb)
        ms = []
        for _ in range(8):
            ctx.render_device(cam, prm)
            st = ctx.sync()
            ms.append(st.gpu_ms)
        line = f"world {world} lanes {lanes} want {nb:2d}: batches {st.batches:2d} device frame {min(ms[2:]):.3f} ms"
        if world == 1:
            e = []
            for _ in range(6):
                t0 = time.perf_counter()
                st = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
                e.append(1e3 * (time.perf_counter() - t0))
            img = pinned.numpy().copy()
            if ref is None:
                ref = img
            line += f" | rt_render (pinned host) {min(e[1:]):.3f} ms, max diff vs 1 lane {np.abs(img - ref).max():.2e}"
        print(line, flush=True)
