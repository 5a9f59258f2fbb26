import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200
from rtb200 import standin
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
ctx.set_pipeline(1, 1); ctx.set_overlap(False); ctx.set_stage_timing(True)
for world in (1, 4, 8, 16, 32, 64, 128, 256):
    ctx.set_shard(0, world)
    best = None
    for _ in range(5):
        ctx.render_device(rtb200.make_camera(), rtb200.make_params(3840, 2160, 0)); st = ctx.sync(); t = ctx.stage_times()
        if best is None or st.gpu_ms < best[0]:
            best = (st.gpu_ms, t, st)
    print(f"1/{world}: frame {best[0]:.3f} ms | extend {best[1]['extend'][0]*1e3:6.1f} us ({best[2].primary_rays} rays) shade {best[1]['shade'][0]*1e3:6.1f} us "
          f"shadow {best[1]['shadow_point'][0]*1e3:6.1f} us ({best[2].shadow_queries} rays) resolve {best[1]['resolve'][0]*1e3:5.1f} us", flush=True)
