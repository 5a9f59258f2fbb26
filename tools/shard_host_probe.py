"""Developer probe: one rank's share of C3 stored straight into a page-locked host image (rt_render_shard) against the same
share left on the device, for worlds of 1 / 2 / 4 / 8 (run on one GPU: one PCIe link)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import torch
import rtb200
from rtb200 import standin
ctx = rtb200.Context(0)
ctx.upload_scene(standin.dragon_standin_scene(), rtb200.BVH_SAH_HOST)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
image = torch.zeros(3840 * 2160 * 3, dtype=torch.float32).pin_memory()
for world in (1, 2, 4, 8):
    ctx.set_shard(0, world)
    dev, host = [], []
    for _ in range(8):
        ctx.render_device(cam, prm)
        dev.append(ctx.sync().gpu_ms)
        host.append(ctx.render_shard_host(cam, prm, image.data_ptr()).gpu_ms)
    mb = 3840 * 2160 * 12 / world / 1e6
    d, h = min(dev[2:]), min(host[2:])
    print(f"1/{world}: device {d:.3f} ms, into the host image {h:.3f} ms (+{h - d:.3f} ms for {mb:.1f} MB = {mb / max(h - d, 1e-6):.1f} GB/s)", flush=True)
