"""Minimal single-GPU driver for ncu: uploads the C3 stand-in scene and renders a few frames (no counters, no CPU work)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = {"sah": rtb200.BVH_SAH_HOST, "lbvh": rtb200.BVH_LBVH_DEVICE}.get(sys.argv[2] if len(sys.argv) > 2 else "", rtb200.BVH_PLOC_DEVICE)  # default: the device builder
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), mode)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
ctx.set_shard(0, int(os.environ.get("RTB200_PROFILE_WORLD", "1")))  # > 1: rank 0's share of the frame split over that many ranks
if os.environ.get("RTB200_ONE_BATCH", "1") == "1":
    ctx.set_pipeline(1, 1)  # one batch per frame: one launch of every kernel per bounce level (ncu serialises anyway)
for _ in range(frames):
    ctx.render_device(cam, prm)
    st = ctx.sync()
print(f"{frames} frames, last {st.gpu_ms:.3f} ms, {st.rays} rays, {st.rays / st.gpu_ms / 1e3:.1f} Mrays/s")
