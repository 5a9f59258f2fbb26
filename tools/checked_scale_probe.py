"""The bounds-checked build at BASELINE sizes: C3 (86 880 triangles, 3840x2160, depth 3) and C5 (8x8x8 lattice, 44.5 M
triangles, 7680x4320, 16 spp, depth 3) rendered device-resident, then the violation counts.  Large scenes and frames are where an index
computed in too narrow a type would leave its array.  Run with RTB200_LIB=.../librtb200_checked.so (tools/checked_run.sh does)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200
from rtb200 import standin

out = {"checked": rtb200.checked_build(), "frames": []}
with rtb200.Context(0) as ctx:
    for name, scene, prm in (
        ("c3", lambda: standin.dragon_standin_scene(), rtb200.make_params(3840, 2160, 3)),
        ("c5", lambda: standin.dragon_lattice_scene(8), rtb200.make_params(7680, 4320, 3, sample_mode=2, sample_size=16)),
    ):
        t0 = time.time()
        sc = scene()
        t1 = time.time()
        ctx.upload_scene(sc, rtb200.BVH_PLOC_DEVICE)
        n_tris = sc.n_tris
        del sc
        t2 = time.time()
        cam = rtb200.make_camera()
        for rep in range(2):  # the second frame of a shape may pick the eight-lanes-per-ray kernels from the first one's queue fills
            ctx.render_device(cam, prm)
            st = ctx.sync()
        out["frames"].append({"config": name, "triangles": n_tris, "rays": st.rays, "gpu_ms": round(st.gpu_ms, 3), "scene_s": round(t1 - t0, 1),
                              "upload_build_s": round(t2 - t1, 2), "violations_so_far": ctx.violations()})
    before = ctx.violations()
    ctx.violations_selftest()
    after = ctx.violations()
    out["selftest"] = {k: after[k] - before[k] for k in after if after[k] != before[k]}
    out["violations"] = before
print(json.dumps(out))
dst = os.environ.get("RTB200_VIOLATIONS_OUT")
if dst:
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
sys.exit(1 if any(out["violations"].values()) else 0)
