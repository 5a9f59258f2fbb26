"""world_size-2 gloo test (CPU) of the multi-rank plumbing bench.py uses: interleaved-tile ownership and the fallback
gather (non-owned pixels stay zero, so a sum-reduce to rank 0 is the gather)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, w, h, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
    import rtb200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owner = rtb200.owner_map(w, h, world)
    yy, xx = np.mgrid[0:h, 0:w]
    image = np.stack([xx, yy, xx * 0 + 7, xx * 0 + 1], -1).astype(np.float32)   # what a full render would hold (rgba)
    fb = np.where((owner == rank)[..., None], image, 0.0).astype(np.float32)     # this rank rendered only its tiles
    t = torch.from_numpy(fb.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    counts = torch.tensor([float((owner == rank).sum())])
    dist.all_reduce(counts)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy().ravel(), image.ravel(), counts.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(96, 48), (70, 45)])
def test_two_rank_gather(tmp_path, w, h):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500) + w
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(2, port, w, h, out), nprocs=2, join=True)
    r = np.load(out)
    n = w * h * 4
    assert np.array_equal(r[:n], r[n:2 * n])
    assert r[-1] == w * h
