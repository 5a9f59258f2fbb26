"""Developer probe: per-launch timeline (RTB200_TRACE_LAUNCHES) of one C3 frame in its real pipelined shape, for rank 0's share of a
frame split over `world` ranks.  usage: RTB200_TRACE_LAUNCHES=1 python tools/timeline_probe.py [world]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
ctx.set_shard(0, world)
for _ in range(6):
    ctx.render_device(cam, prm)
    st = ctx.sync()
print(f"world {world}: plain frame {st.gpu_ms:.3f} ms", file=sys.stderr)
ctx.set_stage_timing(True)
ctx.render_device(cam, prm)
ctx.sync()
sys.stderr.write("---- timeline of the next frame ----\n")
sys.stderr.flush()
os.environ["RTB200_TRACE_LAUNCHES"] = "1"
ctx.render_device(cam, prm)
st = ctx.sync()
print(f"world {world}: frame with events {st.gpu_ms:.3f} ms", file=sys.stderr)
