"""Developer probe: when each band of a host-path C3 frame was rendered and had left (RTB200_TRACE_BANDS=1)."""
import os
import sys
os.environ["RTB200_TRACE_BANDS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import torch
import rtb200
from rtb200 import standin
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
pinned = torch.empty(3840 * 2160 * 3, dtype=torch.float32).pin_memory()
for i in range(4):
    if i == 3:
        sys.stderr.write("---- frame\n"); sys.stderr.flush()
    st = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
sys.stderr.write(f"frame {st.gpu_ms:.3f} ms\n")
