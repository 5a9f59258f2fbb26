"""Developer probe: per-bounce-level stage times of C3, for the whole frame and for one rank's share of a world of N."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam = rtb200.make_camera()
ctx.set_stage_timing(True)
for world in (1, 2, 8):
    ctx.set_shard(0, world)
    prev = {}
    prev_rays = (0, 0, 0)
    spawned = 0
    sec_hist = []
    for depth in range(4):
        prm = rtb200.make_params(3840, 2160, depth)
        best = None
        for _ in range(4):
            ctx.render_device(cam, prm)
            st = ctx.sync()
            t = ctx.stage_times()
            if best is None or st.gpu_ms < best[0]:
                best = (st.gpu_ms, t, (st.primary_rays, st.shadow_queries, st.secondary_rays))
        ms, t, rays = best
        d = {k: t[k][0] - prev.get(k, 0.0) for k in ("extend", "shade", "shadow_point")}
        n_ext = rays[0] if depth == 0 else rays[2] - prev_rays[2]
        dr = (rays[0] + rays[2] - prev_rays[0] - prev_rays[2], rays[1] - prev_rays[1])
        # level `depth` adds: extend over the rays spawned at depth-1 (primary at 0), shade, shadow
        print(f"world {world} level {depth}: frame {ms:7.3f} ms | +extend {d['extend']*1e3:7.1f} us ({n_ext:9d} rays) "
              f"+shade {d['shade']*1e3:6.1f} us +shadow {d['shadow_point']*1e3:7.1f} us ({dr[1]:8d} rays)", flush=True)
        prev = {k: t[k][0] for k in t}
        prev_rays = rays
    print()
