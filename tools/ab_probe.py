"""Developer probe for A/B runs of kernel variants (RTB200_LIB=variants/lib_x.so): C3 frame (median / min of 12), the serialised
per-stage times of one frame, one rank's share of the frame for worlds 2 / 4 / 8, and optionally C2 / C4 / C1 frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402


def frames(ctx, cam, prm, n=12):
    ms = []
    for _ in range(n):
        ctx.render_device(cam, prm)
        ms.append(ctx.sync().gpu_ms)
    ms = ms[2:]
    return float(np.median(ms)), float(min(ms))


def main():
    which = sys.argv[1:] or ["c3"]
    ctx = rtb200.Context(0)
    if os.environ.get("AB_WIDE"):
        ctx.set_wide(int(os.environ["AB_WIDE"]))
    cam = rtb200.make_camera()
    out = [os.path.basename(os.environ.get("RTB200_LIB", "default"))]
    bvh = {"lbvh": rtb200.BVH_LBVH_DEVICE, "ploc": rtb200.BVH_PLOC_DEVICE}.get(os.environ.get("AB_BVH"), rtb200.BVH_SAH_HOST)
    if "c3" in which:
        ctx.upload_scene(standin.dragon_standin_scene(), bvh)
        prm = rtb200.make_params(3840, 2160, 3)
        med, best = frames(ctx, cam, prm)
        out.append(f"C3 {med:.3f}/{best:.3f}")
        ctx.set_pipeline(1, 1)
        ctx.set_overlap(False)
        ctx.set_stage_timing(True)
        ctx.render_device(cam, prm)
        ctx.sync()
        ctx.render_device(cam, prm)
        ctx.sync()
        out.append("stages " + " ".join(f"{k[:6]}={v[0]:.3f}" for k, v in ctx.stage_times().items() if v[1]))
        ctx.set_stage_timing(False)
        ctx.set_pipeline(0, 1)
        ctx.set_overlap(True)
        for world in (2, 4, 8):
            ctx.set_shard(0, world)
            med, best = frames(ctx, cam, prm)
            out.append(f"1/{world} {med:.3f}/{best:.3f}")
        ctx.set_shard(0, 1)
    from util import Golden
    if "c2" in which:
        ctx.upload_scene(Golden("teapot_c2_256x144").scene, bvh)
        med, best = frames(ctx, cam, rtb200.make_params(1920, 1080, 0))
        out.append(f"C2 {med:.3f}/{best:.3f}")
    if "c1" in which:
        ctx.upload_scene(Golden("cornell_c1_256").scene, bvh)
        med, best = frames(ctx, cam, rtb200.make_params(1024, 1024, 3))
        out.append(f"C1 {med:.3f}/{best:.3f}")
    if "c4" in which:
        ctx.upload_scene(Golden("cornell_c4_96").scene, bvh)
        med, best = frames(ctx, cam, rtb200.make_params(2048, 2048, 5, sphere_rays=64), n=6)
        out.append(f"C4 {med:.3f}/{best:.3f}")
    print(" | ".join(out), flush=True)


if __name__ == "__main__":
    main()
