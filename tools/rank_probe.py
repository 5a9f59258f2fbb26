"""Developer probe (torchrun, N ranks): per rank — FP32 issue peak, host time to enqueue one sharded C3 frame, device time of the frame,
and the same with the frame rendered twice back to back (the second enqueue overlaps the first frame: exposes launch-bound ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = rtb200.Context(local)
ctx.upload_scene(standin.dragon_standin_scene())
ctx.set_shard(rank, world)
cam, prm = rtb200.make_camera(), rtb200.make_params(3840, 2160, 3)
peak = ctx.measure_fp32_peak()
for _ in range(5):
    ctx.render_device(cam, prm)
    ctx.sync()
dist.barrier()
enq, dev = [], []
for _ in range(20):
    t0 = time.perf_counter()
    ctx.render_device(cam, prm)
    enq.append(1e3 * (time.perf_counter() - t0))
    dev.append(ctx.sync().gpu_ms)
    dist.barrier()
aff = sorted(os.sched_getaffinity(0))
print(f"rank {rank}: fp32 peak {peak / 1e3:.2f} T/s | enqueue {np.median(enq):.3f} ms | device {np.median(dev):.3f} ms (min {min(dev):.3f}) | cpus {aff[0]}..{aff[-1]} ({len(aff)})", flush=True)
dist.barrier()
dist.destroy_process_group()
