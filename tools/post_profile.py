"""Developer probe: one box-bloom pass on a 4K image (for ncu captures of the post-processing kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402

ctx = rtb200.Context(0)
rng = np.random.default_rng(1)
img = (rng.random((2160, 3840, 3), dtype=np.float32) ** 4 * 6.0).astype(np.float32)
kern = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(3):
    out = ctx.postprocess(img, rtb200.make_post(filtering_option=2, kernel=kern))
print("done", float(out.mean()))
