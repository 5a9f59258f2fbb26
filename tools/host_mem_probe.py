"""Developer probe: rt_render of the C3 frame (wall ms, and the device span of the frame) into three kinds of page-locked host memory —
the CUDA allocator's (rt_host_alloc), ordinary pages registered afterwards (rt_host_register on a numpy array: what a Screen built
on std::vector gets), and registered pages of a 2 MiB-aligned, huge-page-advised mapping — interleaved rounds in one process."""
import ctypes as C
import mmap
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
import rtb200
from rtb200 import standin

W, H = 3840, 2160
nbytes = W * H * 3 * 4
lib = rtb200.lib()
ctx = rtb200.Context(0)
sc = standin.dragon_standin_scene()
ctx.upload_scene(sc, rtb200.BVH_PLOC_DEVICE)
cam, prm = rtb200.make_camera(), rtb200.make_params(W, H, 3)

bufs = {}
p = C.c_void_p()
assert lib.rt_host_alloc(nbytes, C.byref(p)) == 0, lib.rt_last_error()
bufs["cuda allocator"] = (p.value, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (W * H * 3,)))
a = np.zeros(W * H * 3, np.float32)
a[:] = 1.0  # touched, ordinary 4 KiB pages
assert lib.rt_host_register(a.ctypes.data, nbytes) == 0, lib.rt_last_error()
bufs["registered"] = (a.ctypes.data, a)
m = mmap.mmap(-1, nbytes + (4 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
base = C.addressof(C.c_char.from_buffer(m))
aligned = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
try:
    m.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
    print("madvise(MADV_HUGEPAGE):", e)
h = np.ctypeslib.as_array(C.cast(C.c_void_p(aligned), C.POINTER(C.c_float)), (W * H * 3,))
h[:] = 1.0
assert lib.rt_host_register(aligned, nbytes) == 0, lib.rt_last_error()
bufs["registered, huge pages"] = (aligned, h)
try:
    thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    ahp = [l for l in open("/proc/self/smaps_rollup") if "AnonHugePages" in l][0].split()[1]
    print(f"transparent_hugepage: {thp}; AnonHugePages of this process: {ahp} kB")
except Exception as e:
    print("thp state:", e)

res = {k: ([], []) for k in bufs}
frames = {}
for rnd in range(5):
    for k, (ptr, arr) in bufs.items():
        for i in range(6):
            ctx.set_materials(sc.mats)
            ctx.set_lights(sc.point_lights, sc.sphere_lights)
            t0 = time.perf_counter()
            st = ctx.render_host_ptr(cam, prm, ptr)
            t1 = time.perf_counter()
            if rnd and i:
                res[k][0].append(1e3 * (t1 - t0))
                res[k][1].append(st.gpu_ms)
        frames[k] = arr.copy()
for k in bufs:
    assert np.array_equal(frames[k], frames["cuda allocator"]), k
    w, d = res[k]
    print(f"{k:24s} rt_render wall {np.median(w):.3f} / {min(w):.3f} ms (median / min of {len(w)}), device span {np.median(d):.3f} ms", flush=True)
