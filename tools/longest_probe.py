"""Developer probe: the longest query (boxes, triangles) of C3 frames of depth 0..3 (instrumented kernels), whole frame and 1/8 share."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
os.environ["RTB200_TRACE_LAUNCHES"] = "1"
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST if (len(sys.argv) < 2 or sys.argv[1] == "sah") else rtb200.BVH_LBVH_DEVICE)
cam = rtb200.make_camera()
ctx.set_counters(True)
for depth in range(4):
    sys.stderr.write(f"depth {depth}: ")
    sys.stderr.flush()
    ctx.render_device(cam, rtb200.make_params(3840, 2160, depth))
    st = ctx.sync()
    sys.stderr.write(f"   rays {st.rays} boxes/ray {st.node_visits / st.rays:.1f}\n")
