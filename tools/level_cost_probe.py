"""Developer probe: node visits and triangle tests per ray, by bounce level and ray kind (C3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam = rtb200.make_camera()
ctx.set_counters(True)
prev = None
for depth in range(4):
    ctx.render_device(cam, rtb200.make_params(3840, 2160, depth))
    st = ctx.sync()
    cur = dict(ext_rays=st.primary_rays + st.secondary_rays, sh_rays=st.shadow_queries, ext_nodes=st.extend_node_visits, all_nodes=st.node_visits,
               ext_tris=st.extend_tri_tests, all_tris=st.tri_tests)
    d = {k: cur[k] - (prev[k] if prev else 0) for k in cur}
    sh_nodes, sh_tris = d["all_nodes"] - d["ext_nodes"], d["all_tris"] - d["ext_tris"]
    print(f"level {depth}: extend {d['ext_rays']:9d} rays {d['ext_nodes'] / max(1, d['ext_rays']):7.1f} boxes/ray {d['ext_tris'] / max(1, d['ext_rays']):5.2f} tris/ray | "
          f"shadow {d['sh_rays']:9d} rays {sh_nodes / max(1, d['sh_rays']):7.1f} boxes/ray {sh_tris / max(1, d['sh_rays']):5.2f} tris/ray", flush=True)
    prev = cur
