"""Developer probe: small single-GPU frames (depth 3) traced per level and as whole paths (rt_set_paths)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402
from util import Golden  # noqa: E402

ctx = rtb200.Context(0)
cam = rtb200.make_camera()
for name, sc in (("dragon*", standin.dragon_standin_scene()), ("monkey", Golden("monkey_192").scene), ("teapot", Golden("teapot_d3_128x72").scene)):
    ctx.upload_scene(sc)
    for w, h in ((256, 256), (512, 512), (1024, 1024), (1920, 1080)):
        out = []
        for mode in (0, 1):
            ctx.set_paths(mode)
            ms = []
            for _ in range(14):
                ctx.render_device(cam, rtb200.make_params(w, h, 3))
                ms.append(ctx.sync().gpu_ms)
            out.append(float(np.median(ms[4:])))
        print(f"{name:8s} {w}x{h}: per level {out[0]:.3f} ms, paths {out[1]:.3f} ms", flush=True)
ctx.set_paths(-1)
