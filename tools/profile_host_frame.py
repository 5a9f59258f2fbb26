"""Minimal single-GPU driver for ncu: C3 through rt_render into page-locked host memory (the end-to-end path of bench.py)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import torch
import rtb200
from rtb200 import standin
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
pinned = torch.empty(3840 * 2160 * 3, dtype=torch.float32).pin_memory()
for _ in range(frames):
    st = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
print(f"{frames} frames, last {st.gpu_ms:.3f} ms, {st.rays} rays, {st.kernel_launches} launches, image sum {float(pinned.sum()):.1f}")
