"""Developer probe: cost of the Screen post-processing inside a C3 frame (4K), per setting."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
ctx.set_stage_timing(True)
configs = [("off", None), ("box f5", dict(filtering_option=2)), ("gauss f5", dict(filtering_option=2, kernel=1)), ("box f5 x3", dict(filtering_option=1, kernel_repetitions=3)),
           ("box f16", dict(filtering_option=1, filter_size=16)), ("box f17 (direct)", dict(filtering_option=1, filter_size=17)),
           ("exposure+gamma f2", dict(filtering_option=3, filter_size=2, gamma_correction=True))]
for name, cfg in configs:
    ctx.set_postprocess(rtb200.make_post(**cfg) if cfg else None)
    best = None
    for _ in range(6):
        ctx.render_device(cam, prm)
        st = ctx.sync()
        t = ctx.stage_times()["post"]
        if best is None or st.gpu_ms < best[0]:
            best = (st.gpu_ms, t)
    print(f"{name:20s} frame {best[0]:.3f} ms  post stage {best[1][0]:.3f} ms", flush=True)
