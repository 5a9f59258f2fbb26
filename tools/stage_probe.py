"""Developer probe: C3 stage times with kernels running alone (1 lane, no overlap) and the default frame time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
mode = rtb200.BVH_SAH_HOST if (len(sys.argv) < 2 or sys.argv[1] == "sah") else rtb200.BVH_LBVH_DEVICE
ctx.upload_scene(standin.dragon_standin_scene(), mode)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
ms = []
for _ in range(8):
    ctx.render_device(cam, prm)
    ms.append(ctx.sync().gpu_ms)
ctx.set_pipeline(1, 1)
ctx.set_overlap(False)
ctx.set_stage_timing(True)
best = None
for _ in range(5):
    ctx.render_device(cam, prm)
    st = ctx.sync()
    t = ctx.stage_times()
    if best is None or st.gpu_ms < best[0]:
        best = (st.gpu_ms, t)
print(f"default frame {min(ms[2:]):.3f} ms | isolated: frame {best[0]:.3f} ms " + " ".join(f"{k}={v[0]:.3f}" for k, v in best[1].items() if v[1]), flush=True)
