"""Developer probe: build time, node count, depth, boxes per ray and C3 frame time of the three BVH builders (and C5's build on request)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
sc = standin.dragon_standin_scene()
ref = None
for name, mode in (("sah_host", rtb200.BVH_SAH_HOST), ("lbvh", rtb200.BVH_LBVH_DEVICE), ("ploc", rtb200.BVH_PLOC_DEVICE)):
    ctx.upload_scene(sc, mode)   # first build of a mode pays one-time costs (allocations, module load)
    t0 = time.perf_counter()
    ctx.build_bvh(mode)
    build_ms = 1e3 * (time.perf_counter() - t0)
    nodes, depth = ctx.bvh_info()
    ms = []
    for _ in range(12):
        ctx.render_device(cam, prm)
        ms.append(ctx.sync().gpu_ms)
    ctx.set_counters(True)
    ctx.render_device(cam, prm)
    c = ctx.sync()
    ctx.set_counters(False)
    rgb, ids, t, st = ctx.render(cam, rtb200.make_params(960, 540, 3), want_ids=True)
    same = "" if ref is None else f" ids==sah {np.array_equal(ids, ref[0])} t==sah {np.array_equal(t.view(np.int32), ref[1].view(np.int32))} rays==sah {st.rays == ref[2]}"
    if ref is None:
        ref = (ids, t, st.rays)
    print(f"{name:9s} build {build_ms:8.2f} ms nodes {nodes:8d} depth {depth:3d} | C3 frame {np.median(ms[2:]):.3f} ms | boxes/ray {c.node_visits / c.rays:.2f} tris/ray {c.tri_tests / c.rays:.2f}{same}", flush=True)
if "c5" in sys.argv[1:]:
    sc5 = standin.dragon_lattice_scene(8)
    for name, mode in (("lbvh", rtb200.BVH_LBVH_DEVICE), ("ploc", rtb200.BVH_PLOC_DEVICE)):
        ctx.upload_scene(sc5, mode)
        t0 = time.perf_counter()
        ctx.build_bvh(mode)
        build_ms = 1e3 * (time.perf_counter() - t0)
        nodes, depth = ctx.bvh_info()
        p5 = rtb200.make_params(7680, 4320, 3, sample_mode=2, sample_size=16)
        ms = []
        for _ in range(3):
            ctx.render_device(cam, p5)
            ms.append(ctx.sync().gpu_ms)
        print(f"C5 {name:6s} build {build_ms:9.1f} ms nodes {nodes} depth {depth} | frame {min(ms):.2f} ms", flush=True)
