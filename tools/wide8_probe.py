"""Developer probe (RTB200_WIDE8=1): closest hits of incoherent secondary-like rays through the binary tree (one lane per ray) and through the
8-wide tree (eight lanes per ray): equality of the answers and kernel time as a function of the number of rays."""
import os
import sys

import numpy as np

os.environ["RTB200_WIDE8"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import ctypes as C  # noqa: E402

import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

ctx = rtb200.Context(0)
sc = standin.dragon_standin_scene()
ctx.upload_scene(sc)
rng = np.random.default_rng(5)
N = 1 << 20
tri = rng.integers(0, sc.n_tris, N)
P = sc.pos.reshape(-1, 3, 3)[tri]
Nn = sc.nrm.reshape(-1, 3, 3)[tri].mean(1)
w = rng.dirichlet((1, 1, 1), N).astype(np.float32)
p = (P * w[..., None]).sum(1)
d = rng.standard_normal((N, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
d *= np.sign((d * Nn).sum(1, keepdims=True) + 1e-9)      # into the hemisphere of the normal
rays = np.concatenate([p + 0.01 * d, d], 1).astype(np.float32)


def run(n, mode):
    r = rays[:n]
    ids, t = np.empty(n, np.int32), np.empty(n, np.float32)
    best = 1e9
    for _ in range(4):
        rc = rtb200.lib().rt_intersect(ctx._h, r.ctypes.data, n, mode, ids.ctypes.data, t.ctypes.data)
        assert rc == 0, rtb200.lib().rt_last_error()
        best = min(best, rtb200.lib().rt_last_intersect_ms(ctx._h))
    return ids, t, best


for n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
    a_ids, a_t, a_ms = run(n, 1)
    b_ids, b_t, b_ms = run(n, 2)
    same = np.array_equal(a_ids, b_ids) and np.array_equal(a_t.view(np.int32), b_t.view(np.int32))
    print(f"{n:8d} rays: binary {a_ms * 1e3:8.1f} us, 8-wide x 8 lanes {b_ms * 1e3:8.1f} us, same answers {same}, hit {np.mean(a_ids >= 0):.2f}", flush=True)
