"""Developer probe (not the bench contract): frame times of the BASELINE configs on one GPU, per BVH mode."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402


def run(ctx, name, sc, cam, prm, modes=("lbvh", "sah"), reps=5, counters=True):
    for mode in modes:
        t0 = time.time()
        ctx.upload_scene(sc, rtb200.BVH_LBVH_DEVICE if mode == "lbvh" else rtb200.BVH_SAH_HOST)
        build_s = time.time() - t0
        nodes, depth = ctx.bvh_info()
        ms = []
        for _ in range(reps):
            ctx.render_device(cam, prm)
            st = ctx.sync()
            ms.append(st.gpu_ms)
        best = min(ms[1:]) if reps > 1 else ms[0]
        line = f"{name:28s} {mode:5s} build {build_s*1e3:8.1f} ms nodes {nodes:9d} depth {depth:3d} | frame {best:9.3f} ms  rays {st.rays:12d}  {st.rays/best/1e3:9.1f} Mrays/s (prim {st.primary_rays} shad {st.shadow_queries} sec {st.secondary_rays}) launches {st.kernel_launches} batches {st.batches}"
        ctx.set_stage_timing(True)
        ctx.render_device(cam, prm)
        ctx.sync()
        ctx.set_stage_timing(False)
        line += " | stages ms: " + " ".join(f"{k[:7]}={v[0]:.3f}" for k, v in ctx.stage_times().items() if v[1])
        if counters:
            ctx.set_counters(True)
            ctx.render_device(cam, prm)
            c = ctx.sync()
            ctx.set_counters(False)
            line += f" | per ray: nodes {c.node_visits/c.rays:6.2f} tris {c.tri_tests/c.rays:6.2f} full {c.tri_tests_full/c.rays:6.2f}"
        print(line, flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
    ctx = rtb200.Context(0)
    cam = rtb200.make_camera()
    from util import Golden
    if "c1" in which:
        g = Golden("cornell_c1_256")
        run(ctx, "C1 cornell 1024^2 d3", g.scene, cam, rtb200.make_params(1024, 1024, 3))
    if "c2" in which:
        g = Golden("teapot_c2_256x144")
        run(ctx, "C2 teapot 1920x1080 d0", g.scene, cam, rtb200.make_params(1920, 1080, 0))
    if "c3" in which:
        sc = standin.dragon_standin_scene()
        run(ctx, "C3 dragon* 3840x2160 d3", sc, cam, rtb200.make_params(3840, 2160, 3))
    if "c4" in which:
        g = Golden("cornell_c4_96")
        run(ctx, "C4 cornell 2048^2 sph64 d5", g.scene, cam, rtb200.make_params(2048, 2048, 5, sphere_rays=64))
    if "c5" in which:
        t0 = time.time()
        sc = standin.dragon_lattice_scene(8)
        print("lattice built in", time.time() - t0, "s", sc.n_tris, "tris", flush=True)
        run(ctx, "C5 lattice* 7680x4320 16spp", sc, cam, rtb200.make_params(7680, 4320, 3, sample_mode=2, sample_size=16), modes=("lbvh",), reps=2, counters=False)


if __name__ == "__main__":
    main()
