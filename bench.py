#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ray-tracing path (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c1|c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = one frame of the workload (generate -> extend/shade/shadow per bounce -> resolve, plus the tile gather at
N > 1).  Default workload = BASELINE.json configs[2], the config the metric is quoted on: dragon.obj 3840x2160, one
point light with hard shadows, reflection depth 3 — on the seeded procedural STAND-IN mesh, because the reference's
data/dragon.obj is absent (labelled in `data` and `config`).  Rank 0 prints ONE JSON line.

--impl reference times the reference's own CPU implementation of the path (oracle/_ref: its translation units
compiled verbatim; falls back to the restated port if that library is absent) on the host cores, on a bounded
strided sample of the same frame.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))

METRIC = "Mrays/s (all ray types), dragon.obj 4K"
UNIT = "Mrays/s"

WORKLOADS = {
    # name: (description, width, height, max_level, sphere_rays)
    "c3": ("dragon.obj (procedural STAND-IN, 86 880 tris) 3840x2160, 1 point light hard shadows, reflection depth 3, 1 spp", 3840, 2160, 3, 10),
    "c1": ("CornellBox-Mirror-Rotated.obj 1024x1024, 1 point light, hard shadows, reflection depth 3, 1 spp", 1024, 1024, 3, 10),
    "c2": ("teapot.obj 1920x1080 Phong + hard shadows (depth 0)", 1920, 1080, 0, 10),
    "c4": ("CornellBox-Mirror-Rotated.obj 2048x2048 spherical-light soft shadows, 64 samples per hit, depth 5", 2048, 2048, 5, 64),
    "c5": ("8x8x8 lattice of the dragon STAND-IN flattened to one mesh (44.5 M tris), 7680x4320, multipleRays 16 spp, 1 point light, reflection depth 3", 7680, 4320, 3, 10),
}


def load_workload(name):
    import rtb200
    from rtb200 import standin
    desc, w, h, depth, srays = WORKLOADS[name]
    if name == "c3":
        sc = standin.dragon_standin_scene()
    elif name == "c5":  # BASELINE.json configs[4]
        return desc, standin.dragon_lattice_scene(), rtb200.make_camera(), rtb200.make_params(w, h, depth, srays, sample_mode=2, sample_size=16)
    else:  # geometry of the reference's assets travels inside the golden fixtures (tests/golden/make_golden.py)
        d = np.load(os.path.join(ROOT, "tests", "golden", GOLDEN_OF[name] + ".npz"))
        sc = rtb200.SceneData(d["pos"], d["nrm"], d["mesh_id"], d["mats"], d["point_lights"], d["sphere_lights"])
    return desc, sc, rtb200.make_camera(), rtb200.make_params(w, h, depth, srays)


GOLDEN_OF = {"c1": "cornell_c1_256", "c2": "teapot_c2_256x144", "c4": "cornell_c4_96"}


class _Params:  # what cpu_reference_sample reads of an rt_params
    def __init__(self, w, h, depth, srays):
        self.width, self.height, self.max_reflection_level, self.sphere_light_ray_count = w, h, depth, srays


def load_workload_reference(name):
    """The same scenes for the reference arm, built with numpy alone: that arm must not load the product library."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import standin_np
    desc, w, h, depth, srays = WORKLOADS[name]
    if name == "c5":
        raise SystemExit("bench.py --impl reference: C5 (44.5 M triangles) is out of the CPU reference's reach (about 10 GB and seconds per ray)")
    if name == "c3":
        sc = standin_np.dragon_standin_scene()
    else:
        d = np.load(os.path.join(ROOT, "tests", "golden", GOLDEN_OF[name] + ".npz"))
        sc = standin_np.Scene(d["pos"], d["nrm"], d["mesh_id"], d["mats"], d["point_lights"], d["sphere_lights"])
    k = np.float32(0.01745329251994329576923690768489)  # glm::radians' constant, as rtb200.make_camera applies it (src/main.cpp:413-414)
    cam = {"look_at": (0.0, 0.0, 0.0), "euler": tuple(float(np.float32(v) * k) for v in (20.0, 20.0, 0.0)), "dist": 3.0, "fovy": float(np.float32(50.0) * k)}
    return desc, sc, cam, _Params(w, h, depth, srays)


def BVH_MODES(rtb200):
    """--bvh: ploc = built on the device (Morton order, locally-ordered clustering below, binned SAH on top; rt_ploc.cu; the default),
    lbvh = Karras' Morton hierarchy on the device, sah = binned SAH on the host."""
    return {"ploc": rtb200.BVH_PLOC_DEVICE, "lbvh": rtb200.BVH_LBVH_DEVICE, "sah": rtb200.BVH_SAH_HOST}


def job_config(args, desc, world, gather=None):
    """`config` of the JSON line: the job, identical for both arms (the reference arm runs on the GPU arm's config)."""
    return {"workload": desc, "bvh": args.bvh,
            "sharding": f"interleaved 32x16 tiles over {world} GPU(s), scene replicated",
            "gather": gather or ("none (single GPU)" if world == 1 else "peer_store"),
            "l2": "flushed between timed steps (512 MiB memset outside the event pair)"}


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU (NVML, 5 ms period) from before the warm-up to the end of the timed region;
    mark() is called when the timed region starts, so the samples taken inside it can be told apart."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.t_mark = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def mark(self):
        self.t_mark = time.perf_counter()

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        timed = sorted(v for t, v in self.samples if self.t_mark is not None and t >= self.t_mark)
        s = timed if len(timed) >= 3 else sorted(v for _, v in self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "samples_in_timed_region": len(timed)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one has been summarised."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(sc, cam, prm, target_seconds, threads=0, stride=None):
    """Time the reference's CPU path on a strided pixel subset of the same frame (all host threads by default)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    kind = "reference" if oracle.available("reference") else "port"
    o = oracle.Oracle(kind)
    if threads <= 0:
        # every core this process may use — said explicitly, because torchrun exports OMP_NUM_THREADS=1 to its workers
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def run(step):
        _, _, _, st = o.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, sc.sphere_lights, cam, prm.width, prm.height,
                               max_level=prm.max_reflection_level, sphere_rays=prm.sphere_light_ray_count, use_bvh=True, want_ids=False,
                               want_rgb=False, x_step=step, y_step=step, num_threads=threads)
        return st
    step = stride
    if step is None:
        # calibrate on a very sparse subset, then pick the stride that lands near the target time
        probe_step = max(8, int(round((prm.width * prm.height / 2500.0) ** 0.5)))
        st = run(probe_step)
        rate = st.rays / max(st.seconds, 1e-6)
        rays_per_px = st.rays / max(1, ((prm.width + probe_step - 1) // probe_step) * ((prm.height + probe_step - 1) // probe_step))
        want_px = max(256.0, target_seconds * rate / max(rays_per_px, 1e-9))
        step = max(1, int((prm.width * prm.height / want_px) ** 0.5))
        st = run(step)
        if st.seconds < 0.5 * target_seconds and step > 1:  # the sparse probe under-estimates the rate: refine once
            step = max(1, int(step * (st.seconds / target_seconds) ** 0.5))
    st = run(step)
    nx, ny = (prm.width + step - 1) // step, (prm.height + step - 1) // step
    info = {"value": st.rays / st.seconds / 1e6, "unit": UNIT, "cores": int(st.threads), "kind": kind,
            "sample": f"every {step}th pixel in x and y of the {prm.width}x{prm.height} frame ({nx}x{ny} = {nx * ny} primary rays, {st.rays} rays, "
                      f"{st.seconds:.2f} s; useBVH=true, glossy_ray_count=1)",
            "seconds": st.seconds, "rays": int(st.rays), "stride": step}
    return info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc, sc, cam, prm = load_workload_reference(args.workload)
    # each step is a bounded sample (about 3 s of CPU work) so K + W steps end within a few minutes
    infos = []
    stride = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_sample(sc, cam, prm, target_seconds=3.0, stride=stride)
        stride = info["stride"]  # calibrated once, then the same subset every step
        if i >= args.warmup:
            infos.append(info)
    rays = sum(i["rays"] for i in infos)
    secs = sum(i["seconds"] for i in infos)
    value = rays / secs / 1e6
    last = infos[-1]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / len(infos), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": DATA_LABEL[args.workload in ("c3", "c5")],
            "config": job_config(args, desc, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "native_libraries_loaded": _repo_libraries_loaded()}
    _emit(line)
    return 0


def _repo_libraries_loaded():
    """Shared objects of this repository mapped into the process (the reference arm must show the oracle's only)."""
    try:
        libs = {ln.split()[-1] for ln in open("/proc/self/maps") if ln.rstrip().endswith(".so") and ROOT in ln}
        return sorted(os.path.relpath(p, ROOT) for p in libs)
    except OSError:
        return None


DATA_LABEL = {True: "synthetic (procedural dragon stand-in; data/dragon.obj is absent from the reference tree)", False: "reference asset (from tests/golden)"}


# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import rtb200
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL_DEBUG=VERSION makes NCCL print its version on STDOUT, next to the one JSON line this script owes the driver
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    desc, sc, cam, prm = load_workload(args.workload)
    W, H = prm.width, prm.height
    stream = torch.cuda.Stream()
    ctx = rtb200.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    if args.workload == "c5":
        args.no_cpu_baseline = True   # the CPU reference needs ~10 GB and seconds per ray on this mesh
        if args.bvh == "sah":
            args.bvh = "ploc"         # (a host SAH build of 44.5 M triangles takes half a minute)
    bvh_mode = BVH_MODES(rtb200)[args.bvh]
    t0 = time.perf_counter()
    ctx.upload_scene(sc, bvh_mode)
    upload_ms = 1e3 * (time.perf_counter() - t0)   # scene to the device + first build in this process (module load, first allocations)
    t0 = time.perf_counter()
    ctx.build_bvh(bvh_mode)
    build_ms = 1e3 * (time.perf_counter() - t0)    # the build alone: tree, triangle records in leaf order, tie keys
    ctx.set_shard(rank, world)

    # ---- gather target: rank 0's framebuffer, peer-mapped into the other ranks (stores fused into resolve) ----
    gather = "none (single GPU)"
    target = None  # None: the context's own framebuffer
    fb0 = None
    if world > 1:
        gather = "peer_store"
        handle = [ctx.framebuffer_ipc_handle(W, H) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        ok = torch.ones(1, device="cuda")
        if rank != 0:
            try:
                target = ctx.open_peer_framebuffer(handle[0])
            except Exception as e:  # IPC unavailable (e.g. no shared IPC namespace): fall back to an NCCL reduction
                ok.zero_()
                sys.stderr.write(f"[rank {rank}] peer mapping failed ({e}); falling back to NCCL gather\n")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            gather = "nccl_reduce"
            target = None

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def frame():
        ctx.render_device(cam, prm, target)

    gathered = torch.empty(W * H * 4, dtype=torch.float32, device="cuda") if world > 1 else None

    def nccl_gather():
        # fallback gather: non-owned pixels of every rank's framebuffer stay zero, so a sum to rank 0 is the gather
        ptr, _, _ = ctx.framebuffer()
        gathered.copy_(_as_tensor(torch, ptr, W * H * 4))
        dist.reduce(gathered, dst=0, op=dist.ReduceOp.SUM)

    def finish_step():
        if world > 1 and gather == "nccl_reduce":
            nccl_gather()
        st = ctx.sync()
        if world > 1:
            dist.barrier()
        return st

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(max(args.warmup, 3)):
            frame()
            finish_step()
        # ---- timed region: K frames, device time per frame from CUDA events on the launching stream ----
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.mark()
        wall0 = time.perf_counter()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        stats = []
        for k in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record(stream)
            frame()
            if world > 1 and gather == "nccl_reduce":
                nccl_gather()
            ev[k][1].record(stream)
            st = ctx.sync()
            if world > 1:
                dist.barrier()
            stats.append(st)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall_ms = 1e3 * (time.perf_counter() - wall0)
        clocks = sampler.finish()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if os.environ.get("RTB200_BENCH_VERBOSE"):
        sys.stderr.write(f"[rank {rank}] mean device ms per step {float(step_ms.mean().item()):.3f}, rays per step {stats[-1].rays}\n")
    rays = torch.tensor([float(s.rays) for s in stats], dtype=torch.float64, device="cuda")
    traced = torch.tensor([float(s.traced_rays) for s in stats], dtype=torch.float64, device="cuda")
    gather_bytes = torch.tensor([float(stats[-1].gather_bytes)], dtype=torch.float64, device="cuda")
    launches = sum(s.kernel_launches for s in stats)
    per_rank_ms = [float(step_ms.mean().item())]
    if world > 1:
        mine = torch.tensor(per_rank_ms, dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank_ms = [float(t.item()) for t in every]
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        dist.all_reduce(traced, op=dist.ReduceOp.SUM)
        dist.all_reduce(gather_bytes, op=dist.ReduceOp.SUM)
    total_ms = float(step_ms.sum().item())
    total_rays = float(rays.sum().item())
    value = total_rays / total_ms / 1e3  # Mrays/s
    traced_value = float(traced.sum().item()) / total_ms / 1e3

    # ---- end to end through the reference-facing call: host buffers, copies inside the timed region ----
    pinned = torch.empty(H * W * 3, dtype=torch.float32).pin_memory()
    h2d_bytes = sc.mats.nbytes + sc.point_lights.nbytes + sc.sphere_lights.nbytes + 32 + 36  # materials + lights + camera + params
    e2e_steps = max(3, min(args.steps, 10))

    # N > 1: the host image is one shared-memory segment that every rank registers with CUDA (page-locked, mapped); each rank
    # stores the tiles it owns straight into it over its own PCIe link (rt_render_shard) and a barrier ends the step.  If the
    # segment cannot be shared or registered, rank 0 downloads the gathered device frame instead.
    shared, shared_ptr, e2e_path = None, 0, "rt_render into page-locked memory (kernels store the frame themselves: background rows during the frame, rows with hits when their batch is resolved)"
    if world > 1:
        from multiprocessing import shared_memory
        nbytes = H * W * 3 * 4
        ok = torch.ones(1, device="cuda")
        name = [None]
        if rank == 0:
            try:
                vfs = os.statvfs("/dev/shm")   # a segment larger than the tmpfs would end in SIGBUS on first touch, not in an exception
                if vfs.f_bavail * vfs.f_frsize < nbytes + (16 << 20):
                    raise RuntimeError("/dev/shm is too small for the frame")
                shared = shared_memory.SharedMemory(create=True, size=nbytes)
                name[0] = shared.name
            except Exception as e:
                sys.stderr.write(f"[rank 0] no shared-memory segment ({e})\n")
        dist.broadcast_object_list(name, src=0)    # every rank gets here, whatever happened on rank 0
        try:
            if name[0] is None:
                raise RuntimeError("rank 0 could not create the segment")
            if rank != 0:
                shared = shared_memory.SharedMemory(name=name[0])
                try:  # the creator unlinks it; Python < 3.13 would have every attaching process try as well
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(shared._name, "shared_memory")
                except Exception:
                    pass
            view = np.ndarray((H * W * 3,), dtype=np.float32, buffer=shared.buf)
            if rank == 0:
                view[:] = 0.0
            shared_ptr = view.ctypes.data
            if rtb200.lib().rt_host_register(shared_ptr, nbytes) != 0:  # page-locked, mapped into the device
                raise RuntimeError("rt_host_register: " + rtb200.lib().rt_last_error().decode())
        except Exception as e:
            ok.zero_()
            sys.stderr.write(f"[rank {rank}] shared host image unavailable ({e}); rank 0 downloads the gathered frame\n")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            shared_ptr = 0
        e2e_path = "rt_render_shard (every rank stores its tiles into one shared page-locked host image)" if shared_ptr else "rt_render_device + gather + rank 0 downloads"

    host_store_gbs = None
    if world > 1 and shared_ptr:
        # What one rank's link carries while ALL ranks store into host memory (GPUs behind one PCIe switch share its uplink; the
        # host's ingest rate is finite): measured, not assumed — every rank copies 64 MiB device-to-host at the same moment, 6 times.
        probe_d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        probe_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        rates = []
        for k in range(6):
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            probe_h.copy_(probe_d, non_blocking=True)
            b.record()
            b.synchronize()
            if k >= 2:
                rates.append(probe_d.numel() / (a.elapsed_time(b) * 1e6))
        host_store_gbs = float(np.median(rates))
        ctx.set_host_store_rate(host_store_gbs)
        del probe_d, probe_h

    def e2e_step():
        ctx.set_materials(sc.mats)                      # the reference re-reads materials and lights every frame
        ctx.set_lights(sc.point_lights, sc.sphere_lights)
        if world == 1:
            return ctx.render_host_ptr(cam, prm, pinned.data_ptr())
        if shared_ptr:
            st = ctx.render_shard_host(cam, prm, shared_ptr)
            dist.barrier()
            return st
        ctx.render_device(cam, prm, target)
        st = finish_step()
        if rank == 0:
            ptr = gathered.data_ptr() if gather == "nccl_reduce" else ctx.framebuffer()[0]
            ctx.download_rgb_ptr(ptr, W, H, pinned.data_ptr())
        return st
    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    e_rays = 0.0
    for _ in range(e2e_steps):
        e_rays += e2e_step().rays
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_ms = torch.tensor([1e3 * (time.perf_counter() - e0)], dtype=torch.float64, device="cuda")
    e_r = torch.tensor([e_rays], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e_r, op=dist.ReduceOp.SUM)
    e2e_value = float(e_r.item()) / float(e_ms.item()) / 1e3
    if world > 1 and shared_ptr:
        # the shared host image must be the frame the device-side gather produces
        ctx.render_device(cam, prm, target)
        finish_step()
        if rank == 0:
            ctx.download_rgb_ptr(gathered.data_ptr() if gather == "nccl_reduce" else ctx.framebuffer()[0], W, H, pinned.data_ptr())
            diff = float(np.abs(np.ndarray((H * W * 3,), dtype=np.float32, buffer=shared.buf) - pinned.numpy()).max())
            if not diff <= 1e-6:
                raise RuntimeError(f"shared host image differs from the gathered frame by {diff}")
        dist.barrier()
        rtb200.lib().rt_host_unregister(shared_ptr)
    if shared is not None:
        shared.close()
        if rank == 0:
            shared.unlink()
    gather_check = None
    if world > 1:
        # the frame the ranks gathered must be the frame ONE GPU renders: rank 0 renders it unsharded, once, untimed
        ctx.render_device(cam, prm, target)
        finish_step()
        if rank == 0:
            got = np.empty(H * W * 3, np.float32)
            ctx.download_rgb_ptr(gathered.data_ptr() if gather == "nccl_reduce" else ctx.framebuffer()[0], W, H, got.ctypes.data)
            ctx.set_shard(0, 1)
            ctx.render_device(cam, prm)
            ctx.sync()
            want = np.empty(H * W * 3, np.float32)
            ctx.download_rgb_ptr(ctx.framebuffer()[0], W, H, want.ctypes.data)
            ctx.set_shard(rank, world)
            diff = float(np.abs(got - want).max())
            lit = int(np.count_nonzero(want))
            if not diff <= 1e-6 or lit == 0:
                raise RuntimeError(f"the gathered frame differs from the unsharded frame by {diff} ({lit} non-zero values)")
            gather_check = {"max_abs_diff_vs_unsharded_frame": diff, "non_zero_values": lit}
        dist.barrier()

    # ---- roofline of the dominant kernel (separate, untimed passes: stage events, then instrumented counters) ----
    fp32_peak = ctx.measure_fp32_peak()   # un-fused FMUL / FADD issue rate of this very GPU (microbenchmark, untimed)
    stage, c = stage_and_counter_passes(ctx, frame, finish_step, flush)
    roof = roofline(stage, c, args.workload, fp32_peak)

    # ---- the other configs of BASELINE.json, untimed side measurements: C1 / C2 / C4 on one GPU, C5 on eight ----
    extra_configs = {}
    if world > 1 and target:   # the headline's gather target is no longer needed (C5 below exports its own, larger one)
        ctx.close_peer_framebuffer(target)
        target = None
    if world > 1:
        dist.barrier()
    if not args.no_extra_configs and args.workload == "c3":
        if world == 1:
            for name in ("c1", "c2", "c4"):
                extra_configs[name] = measure_config(ctx, name, flush, fp32_peak, None)
        elif world == 8 and not os.environ.get("RTB200_BENCH_SKIP_C5"):
            extra_configs["c5"] = measure_config(ctx, "c5", flush, fp32_peak, dict(dist=dist, rank=rank, world=world, torch=torch))

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rdesc, rsc, rcam, rprm = load_workload_reference(args.workload)
            cpu = cpu_reference_sample(rsc, rcam, rprm, target_seconds=15.0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        st = stats[-1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": DATA_LABEL[args.workload in ("c3", "c5")],
            "config": job_config(args, desc, world, gather),
            "clocks": clocks, "wall_ms_per_step_incl_flush_and_sync": wall_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes) * world, "d2h_bytes_per_step": int(W * H * 3 * 4), "steps": e2e_steps,
                    "ms_per_step": float(e_ms.item()) / e2e_steps, "path": e2e_path,
                    "host_store_gbs_per_rank_all_ranks_copying": host_store_gbs},
            "gpu_launches": int(launches),
            "roofline": roof,
            "extra": {
                "bvh_build_ms": build_ms, "scene_upload_and_first_build_ms": upload_ms,
                "rays_per_frame": {"primary": int(st.primary_rays), "shadow": int(st.shadow_queries), "secondary": int(st.secondary_rays),
                                   "primary_traced": int(st.traced_primary_rays)} if world == 1 else {"all_ranks": int(total_rays / args.steps)},
                # `value` counts every ray the reference casts (its accounting: each primary ray, each cansee iteration, each reflection / refraction ray);
                # primary rays of pixels outside the projection of the scene's bounding box are answered as misses without walking the BVH:
                "traced_mrays_per_s": traced_value, "traced_fraction_of_counted_rays": traced_value / value,
                "per_rank_ms_per_step": per_rank_ms,
                "gather_bytes_per_frame": int(gather_bytes.item()),
                "gather_check": gather_check,
                "configs": extra_configs,
            },
        }
        if cpu:
            line["cpu_baseline"] = cpu
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def stage_and_counter_passes(ctx, frame, finish_step, flush):
    """Per-stage device times (every kernel alone: one batch, one stream, events around each launch; median of 3 frames) and the
    instrumented counters of one more frame."""
    ctx.set_pipeline(1, 1)      # one batch, one stream: every kernel runs alone, so its events time it in isolation
    ctx.set_overlap(False)
    ctx.set_stage_timing(True)
    stage_ms = {}
    for _ in range(3):
        flush.zero_()
        frame()
        finish_step()
        for k, (ms, n) in ctx.stage_times().items():
            stage_ms.setdefault(k, []).append((ms, n))
    ctx.set_stage_timing(False)
    stage = {k: (float(np.median([m for m, _ in v])), v[0][1]) for k, v in stage_ms.items()}
    ctx.set_counters(True)
    frame()
    c = finish_step()
    ctx.set_counters(False)
    ctx.set_pipeline(0, 1)
    ctx.set_overlap(True)
    return stage, c


def measure_config(ctx, name, flush, fp32_peak, mg):
    """One of the other BASELINE.json configs, measured beside the headline (untimed side measurement; never the bench value):
    frame time (median of 5, L2 flushed), rays/s counted and traced, and both roofline fractions of its dominant kernel.
    mg: None on one GPU; the torch.distributed plumbing when the frame is sharded (C5 on eight GPUs, gathered on rank 0)."""
    import rtb200
    t0 = time.perf_counter()
    try:
        # set-up: the part that can fail on one rank alone (44.5 M triangles: 3 GB of host arrays and 12 GB of build scratch per rank).
        # The ranks agree on the outcome before any of them enters a collective of the measurement itself.
        target, setup_error = None, None
        try:
            desc, sc, cam, prm = load_workload(name)
            ctx.upload_scene(sc, rtb200.BVH_PLOC_DEVICE)
            del sc
        except Exception as e:
            setup_error = f"{type(e).__name__}: {e}"
        if mg:
            dist, torch = mg["dist"], mg["torch"]
            ok = torch.tensor([0.0 if setup_error else 1.0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                return {"error": setup_error or "set-up failed on another rank"}
            handle = [ctx.framebuffer_ipc_handle(prm.width, prm.height) if mg["rank"] == 0 else None]
            dist.broadcast_object_list(handle, src=0)
            if mg["rank"] != 0:
                target = ctx.open_peer_framebuffer(handle[0])
            dist.barrier()
        elif setup_error:
            return {"error": setup_error}

        def frame():
            ctx.render_device(cam, prm, target)

        def finish():
            st = ctx.sync()
            if mg:
                mg["dist"].barrier()
            return st
        for _ in range(2):
            frame()
            finish()
        ms, st = [], None
        for _ in range(5):
            flush.zero_()
            frame()
            st = finish()
            ms.append(st.gpu_ms)
        t = float(np.median(ms))
        rays, traced = float(st.rays), float(st.traced_rays)
        if mg:
            v = mg["torch"].tensor([t, rays, traced], dtype=mg["torch"].float64, device="cuda")
            vmax = v.clone()
            mg["dist"].all_reduce(vmax, op=mg["dist"].ReduceOp.MAX)
            mg["dist"].all_reduce(v, op=mg["dist"].ReduceOp.SUM)
            t, rays, traced = float(vmax[0].item()), float(v[1].item()), float(v[2].item())
        stage, c = stage_and_counter_passes(ctx, frame, finish, flush)
        roof = roofline(stage, c, name, fp32_peak)
        if target:
            ctx.close_peer_framebuffer(target)
        return {"workload": desc, "n_gpus": mg["world"] if mg else 1, "ms_per_frame": t, "mrays_per_s_counted": rays / t / 1e3, "mrays_per_s_traced": traced / t / 1e3,
                "kernel": roof["kernel"], "fp32_issue_frac": roof["frac"], "hbm_frac_algorithmic": roof["hbm"]["frac_algorithmic"],
                "per_ray": roof["per_ray"], "stage_ms": roof["stage_ms"], "seconds_spent": time.perf_counter() - t0,
                "note": "rank 0's share of the sharded frame for the per-kernel figures" if mg else None}
    except Exception as e:  # a side measurement must not cost the headline line
        return {"error": f"{type(e).__name__}: {e}"}


def _as_tensor(torch, ptr, n_floats):
    """View raw device memory owned by the C library as a torch tensor (for NCCL plumbing only)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")


def roofline(stage, c, workload, fp32_peak_ginst):
    """Roofline of the dominant kernel.  What binds this path is FP32 issue / latency under divergence, not HBM: the scene (BVH +
    triangles, about 11 MB for C1-C4) is L2-resident, so `achieved` / `peak` / `frac` are ALGORITHMIC flops per second against the
    un-fused FMUL / FADD issue rate measured on this GPU in this run; the HBM figures (algorithmic bytes, which L1 / L2 serve, and the
    DRAM traffic ncu measured) are reported beside it.  Algorithmic work per ray follows SURVEY section 8(d): 32 B per BVH node fetched +
    64 B per triangle fetched + the ray's own record traffic; 24 flops per box test, 12 per plane stage, 57 per full triangle stage
    (+ fixed part).  Node / triangle counts are measured by the instrumented kernels."""
    peak, peak_src = measured_peaks()
    ext_rays = c.primary_rays + c.secondary_rays
    sh_rays = c.shadow_queries
    sh_nodes, sh_tris, sh_full = c.node_visits - c.extend_node_visits, c.tri_tests - c.extend_tri_tests, c.tri_tests_full - c.extend_tri_tests_full
    kernels = {
        "extend": dict(ms=stage["extend"][0], launches=stage["extend"][1], rays=ext_rays,
                       bytes=c.extend_node_visits * 32 + c.extend_tri_tests * 64 + ext_rays * (32 + 8),
                       flops=c.extend_node_visits * 24 + c.extend_tri_tests * 12 + c.extend_tri_tests_full * 57 + ext_rays * 20),
        "shadow_point": dict(ms=stage["shadow_point"][0], launches=stage["shadow_point"][1], rays=sh_rays,
                             bytes=sh_nodes * 32 + sh_tris * 64 + sh_rays * (48 + 12),
                             flops=sh_nodes * 24 + sh_tris * 12 + sh_full * 57 + sh_rays * 30),
    }
    if stage.get("shadow_sphere", (0, 0))[1]:
        kernels["shadow_sphere"] = dict(kernels.pop("shadow_point"), ms=stage["shadow_sphere"][0], launches=stage["shadow_sphere"][1])
    name = max(kernels, key=lambda k: kernels[k]["ms"])
    k = kernels[name]
    total_ms = sum(v[0] for v in stage.values())
    launch_ms = k["ms"] / max(1, k["launches"])
    tflops = k["flops"] / max(k["ms"], 1e-9) / 1e9      # algorithmic FP32 operations per second, in 1e12
    peak_t = fp32_peak_ginst / 1e3
    gbs = k["bytes"] / max(k["ms"], 1e-9) / 1e6          # algorithmic GB/s
    traffic = ncu_traffic().get(name) if workload == "c3" else None  # the committed ncu capture is of the C3 frame
    nominal_t = 148 * 128 * 1.965e9 / 1e12
    return {"bound": "fp32_issue", "kernel": "k_" + name, "achieved": tflops, "peak": peak_t, "unit": "TFLOP/s", "frac": tflops / max(peak_t, 1e-9), "traffic": traffic,
            "peak_source": "measured in this run: un-fused FMUL/FADD issue-rate microbenchmark (rt_measure_fp32_peak); nominal 148 SM x 128 lanes x 1.965 GHz = %.1f T/s" % nominal_t,
            "frac_of_nominal_issue": tflops / nominal_t,
            "launches_per_step": k["launches"], "avg_launch_ms": launch_ms,
            "algorithmic_flops_per_launch": k["flops"] / max(1, k["launches"]), "algorithmic_bytes_per_launch": k["bytes"] / max(1, k["launches"]),
            "share_of_step": k["ms"] / max(total_ms, 1e-9),
            "per_ray": {"nodes": (c.node_visits / max(1, c.rays)), "tris": c.tri_tests / max(1, c.rays), "tris_full": c.tri_tests_full / max(1, c.rays)},
            "stage_ms": {s: round(v[0], 4) for s, v in stage.items()},
            "hbm": {"achieved_algorithmic_gbs": gbs, "peak_gbs": peak, "frac_algorithmic": gbs / peak, "peak_source": peak_src,
                    "dram_gbs_measured": (traffic / (launch_ms * 1e-3) / 1e9) if traffic else None,
                    "frac_measured": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                    "note": "algorithmic node / triangle bytes are served by L1 / L2 (scene ~11 MB, L2-resident); `traffic` = DRAM bytes per launch from the committed ncu capture"}}


_REAL_STDOUT = None


def _own_stdout():
    """Libraries below (NCCL prints its version line) write to file descriptor 1; the driver expects ONE JSON line there.  Point fd 1
    at stderr for the rest of the run and keep the real stdout for that line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--bvh", default="ploc", choices=["ploc", "sah", "lbvh"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the side measurements of C1 / C2 / C4 (one GPU) and C5 (eight GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 20 and "--steps" not in " ".join(sys.argv):
            args.steps = 3
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
