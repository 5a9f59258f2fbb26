"""ctypes binding of librtb200.so (include/rt_b200.h) for tests, bench.py and torch.distributed plumbing.

The product is the shared library; this module only marshals numpy / torch buffers into its C ABI.  There is no
CPU fallback: importing works without a GPU (the library loads), but every compute entry point needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(os.path.dirname(_HERE), "librtb200.so")  # override: kernel-variant experiments

RT_OK, RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_OVERFLOW, RT_ERR_IO = 0, 1, 2, 3, 4
BVH_LBVH_DEVICE, BVH_SAH_HOST, BVH_AUTO, BVH_PLOC_DEVICE = 0, 1, 2, 3


class RtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rt_b200 error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("ks", C.c_float * 3), ("shininess", C.c_float), ("transparency", C.c_float)]


class PointLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3)]


class SphereLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("radius", C.c_float), ("color", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("look_at", C.c_float * 3), ("euler", C.c_float * 3), ("dist", C.c_float), ("fovy", C.c_float)]


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("max_reflection_level", C.c_int), ("sphere_light_ray_count", C.c_int),
                ("glossy_ray_count", C.c_int), ("refraction_factor", C.c_float), ("sample_mode", C.c_int), ("sample_size", C.c_int),
                ("use_bvh", C.c_int), ("exhaustive", C.c_int), ("plane_light_ray_count_1d", C.c_int), ("texture_debug", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_queries", C.c_uint64), ("secondary_rays", C.c_uint64), ("node_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("tri_tests_full", C.c_uint64), ("gpu_ms", C.c_float), ("kernel_launches", C.c_int),
                ("batches", C.c_int), ("extend_node_visits", C.c_uint64), ("extend_tri_tests", C.c_uint64), ("extend_tri_tests_full", C.c_uint64),
                ("traced_primary_rays", C.c_uint64), ("gather_bytes", C.c_uint64)]

    @property
    def traced_rays(self) -> int:
        """Rays that walked the BVH: everything but the primary rays of pixels outside the scene's projection."""
        return int(self.traced_primary_rays + self.shadow_queries + self.secondary_rays)

    @property
    def rays(self) -> int:
        return int(self.primary_rays + self.shadow_queries + self.secondary_rays)


MATERIAL_DTYPE = np.dtype([("kd", np.float32, 3), ("ks", np.float32, 3), ("shininess", np.float32), ("transparency", np.float32)])

_lib = None

# every symbol include/rt_b200.h declares: (name, restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.POINTER(C.c_float)
class Texture(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("rgb", C.c_void_p)]


class TextureParams(C.Structure):
    _fields_ = [("filtering", C.c_int), ("out_of_bounds_x", C.c_int), ("out_of_bounds_y", C.c_int), ("border_color", C.c_float * 3)]


TEX_NEAREST, TEX_BILINEAR, TEX_MIP_NEAREST, TEX_MIP_BILINEAR, TEX_TRILINEAR = range(5)
OOB_BORDER, OOB_CLAMP, OOB_REPEAT = 0, 1, 2


class PostParams(C.Structure):
    """rt_post_params: Screen's bloom / tone-mapping / gamma settings (include/rt_b200.h)."""
    _fields_ = [("filtering_option", C.c_int), ("kernel", C.c_int), ("kernel_repetitions", C.c_int), ("filter_size", C.c_int),
                ("sigma", C.c_float), ("exposure", C.c_float), ("gamma_correction", C.c_int), ("gamma", C.c_float), ("bloom_live", C.c_int)]


FILTER_NONE, FILTER_BLOOM, FILTER_BLOOM_REINHARD, FILTER_BLOOM_EXPOSURE, FILTER_ONLY_LIGHT, FILTER_ONLY_LIGHT_KERNEL = range(6)
KERNEL_BOX, KERNEL_GAUSSIAN = 0, 1


def make_post(filtering_option=FILTER_NONE, kernel=KERNEL_BOX, kernel_repetitions=1, filter_size=5, sigma=2.0, exposure=0.5,
              gamma_correction=False, gamma=2.2, bloom_live=True) -> PostParams:
    """Defaults are Screen's (src/screen.h:84-101), except bloom_live, which postprocessImage needs to bloom at all."""
    return PostParams(int(filtering_option), int(kernel), int(kernel_repetitions), int(filter_size), float(sigma), float(exposure),
                      1 if gamma_correction else 0, float(gamma), 1 if bloom_live else 0)


SYMBOLS = [
    ("rt_create", _I, [_I, C.POINTER(_P)]),
    ("rt_destroy", _I, [_P]),
    ("rt_set_stream", _I, [_P, _P]),
    ("rt_upload_scene", _I, [_P, _P, _P, _P, C.c_int64, _P, _I]),
    ("rt_build_bvh", _I, [_P, _I]),
    ("rt_set_materials", _I, [_P, _P, _I]),
    ("rt_set_lights", _I, [_P, _P, _I, _P, _I]),
    ("rt_set_spheres", _I, [_P, _P, _I]),
    ("rt_set_spot_lights", _I, [_P, _P, _I]),
    ("rt_set_plane_lights", _I, [_P, _P, _I]),
    ("rt_set_texcoords", _I, [_P, _P]),
    ("rt_set_textures", _I, [_P, C.POINTER(Texture), _I, _P, _I]),
    ("rt_set_texturing", _I, [_P, C.POINTER(TextureParams)]),
    ("rt_set_postprocess", _I, [_P, C.POINTER(PostParams)]),
    ("rt_postprocess", _I, [_P, C.POINTER(PostParams), _P, _I, _I, _I, _P]),
    ("rt_postprocess_device", _I, [_P, C.POINTER(PostParams), _P, _I, _I]),
    ("rt_bvh_info", _I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    ("rt_set_counters", _I, [_P, _I]),
    ("rt_set_batch_rays", _I, [_P, C.c_uint]),
    ("rt_set_overlap", _I, [_P, _I]),
    ("rt_measure_fp32_peak", _I, [_P, C.POINTER(C.c_double)]),
    ("rt_set_paths", _I, [_P, _I]),
    ("rt_set_wide", _I, [_P, _I]),
    ("rt_set_pipeline", _I, [_P, _I, _I, C.c_uint]),
    ("rt_set_stage_timing", _I, [_P, _I]),
    ("rt_stage_times", _I, [_P, _P, _P]),
    ("rt_set_shard", _I, [_P, _I, _I]),
    ("rt_render", _I, [_P, C.POINTER(Camera), C.POINTER(Params), _P, _P, _P, C.POINTER(Stats)]),
    ("rt_render_device", _I, [_P, C.POINTER(Camera), C.POINTER(Params), _P]),
    ("rt_visible_rect", _I, [C.POINTER(Camera), C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    ("rt_render_shard", _I, [_P, C.POINTER(Camera), C.POINTER(Params), _P, C.POINTER(Stats)]),
    ("rt_sync", _I, [_P, C.POINTER(Stats)]),
    ("rt_framebuffer", _I, [_P, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I)]),
    ("rt_framebuffer_ipc_handle", _I, [_P, _I, _I, _P]),
    ("rt_open_peer_framebuffer", _I, [_P, _P, C.POINTER(_P)]),
    ("rt_set_gather_target", _I, [_P, _P]),
    ("rt_set_host_store_rate", _I, [_P, C.c_double]),
    ("rt_host_register", _I, [_P, C.c_size_t]),
    ("rt_host_unregister", _I, [_P]),
    ("rt_host_alloc", _I, [C.c_size_t, C.POINTER(_P)]),
    ("rt_host_free", _I, [_P]),
    ("rt_current_device", _I, []),
    ("rt_close_peer_framebuffer", _I, [_P, _P]),
    ("rt_download_rgb", _I, [_P, _P, _I, _I, _P]),
    ("rt_intersect", _I, [_P, _P, C.c_int64, _I, _P, _P]),
    ("rt_last_intersect_ms", C.c_float, [_P]),
    ("rt_checked_build", _I, []),
    ("rt_violations", _I, [_P, C.POINTER(C.c_uint), _I]),
    ("rt_violations_selftest", _I, [_P]),
    ("rt_load_obj", _I, [C.c_char_p, _I, C.POINTER(_P)]),
    ("rt_soup_num_triangles", C.c_int64, [_P]),
    ("rt_soup_num_meshes", _I, [_P]),
    ("rt_soup_positions", _P, [_P]),
    ("rt_soup_normals", _P, [_P]),
    ("rt_soup_mesh_ids", _P, [_P]),
    ("rt_soup_texcoords", _P, [_P]),
    ("rt_soup_num_textures", _I, [_P]),
    ("rt_soup_textures", _P, [_P]),
    ("rt_soup_material_textures", _P, [_P]),
    ("rt_soup_materials", _P, [_P]),
    ("rt_soup_free", None, [_P]),
    ("rt_last_error", C.c_char_p, []),
    ("rt_version", C.c_char_p, []),
]


def lib() -> C.CDLL:
    """Load librtb200.so; raises (never falls back) if the CUDA extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C raytracer-group27_b200` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc: int) -> None:
    if rc != RT_OK:
        raise RtError(rc, lib().rt_last_error().decode(errors="replace"))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class SceneData:
    """Triangle soup in the reference's global triangle order + per-mesh materials + lights."""
    pos: np.ndarray          # (n, 9) float32
    nrm: np.ndarray          # (n, 9) float32
    mesh_id: np.ndarray      # (n,) int32
    mats: np.ndarray         # (m,) MATERIAL_DTYPE
    point_lights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))   # position, colour
    sphere_lights: np.ndarray = field(default_factory=lambda: np.zeros((0, 7), np.float32))  # position, radius, colour
    spheres: np.ndarray = field(default_factory=lambda: np.zeros((0, 12), np.float32))       # centre, radius, kd, ks, shininess, transparency
    spot_lights: np.ndarray = field(default_factory=lambda: np.zeros((0, 10), np.float32))   # position, direction, angle (deg), colour
    plane_lights: np.ndarray = field(default_factory=lambda: np.zeros((0, 12), np.float32))  # position, width, height, colour
    uv: "np.ndarray | None" = None           # (n, 6) float32: texture coordinates of the three corners
    textures: list = field(default_factory=list)   # (H, W, 3) uint8 images (top row first), as stbi_load delivers them
    mesh_tex: "np.ndarray | None" = None     # (m,) int32: texture of each mesh, -1 = none

    @property
    def n_tris(self) -> int:
        return int(self.pos.shape[0])


def load_obj(path: str, normalize: bool = False) -> SceneData:
    """loadMesh (src/mesh.cpp:58-188) through the library's host-side OBJ/MTL importer (no GPU needed)."""
    l = lib()
    h = _P()
    _check(l.rt_load_obj(os.fsencode(path), 1 if normalize else 0, C.byref(h)))
    try:
        n = l.rt_soup_num_triangles(h)
        m = l.rt_soup_num_meshes(h)
        pos = np.ctypeslib.as_array(C.cast(l.rt_soup_positions(h), _F), shape=(n, 9)).copy()
        nrm = np.ctypeslib.as_array(C.cast(l.rt_soup_normals(h), _F), shape=(n, 9)).copy()
        ids = np.ctypeslib.as_array(C.cast(l.rt_soup_mesh_ids(h), C.POINTER(C.c_int)), shape=(n,)).copy()
        raw = np.ctypeslib.as_array(C.cast(l.rt_soup_materials(h), _F), shape=(m, 8)).copy()
        mats = np.zeros(m, MATERIAL_DTYPE)
        mats["kd"], mats["ks"], mats["shininess"], mats["transparency"] = raw[:, 0:3], raw[:, 3:6], raw[:, 6], raw[:, 7]
        uv = np.ctypeslib.as_array(C.cast(l.rt_soup_texcoords(h), _F), shape=(n, 6)).copy()
        nt = l.rt_soup_num_textures(h)
        mesh_tex = np.ctypeslib.as_array(C.cast(l.rt_soup_material_textures(h), C.POINTER(C.c_int)), shape=(m,)).copy().astype(np.int32)
        textures = []
        tex_arr = C.cast(l.rt_soup_textures(h), C.POINTER(Texture))
        for k in range(nt):
            t = tex_arr[k]
            texels = np.ctypeslib.as_array(C.cast(t.rgb, _F), shape=(t.height, t.width, 3))
            textures.append(np.rint(texels * 255.0).astype(np.uint8))   # texels are byte / 255.0f: back to the file's bytes
    finally:
        l.rt_soup_free(h)
    sc = SceneData(pos, nrm, ids.astype(np.int32), mats)
    sc.uv, sc.textures, sc.mesh_tex = uv, textures, mesh_tex
    return sc


def make_camera(look_at=(0.0, 0.0, 0.0), euler_deg=(20.0, 20.0, 0.0), dist=3.0, fovy_deg=50.0) -> Camera:
    """Reference default camera (src/main.cpp:413-414); degrees -> radians with glm::radians' float constant."""
    k = np.float32(0.01745329251994329576923690768489)
    cam = Camera()
    cam.look_at[:] = [float(v) for v in look_at]
    cam.euler[:] = [float(np.float32(v) * k) for v in euler_deg]
    cam.dist = float(dist)
    cam.fovy = float(np.float32(fovy_deg) * k)
    return cam


def visible_rect(cam: Camera, width: int, height: int, lo, hi):
    """rt_visible_rect: (x0, x1, y0, y1), half-open, y from the bottom."""
    l = (C.c_float * 3)(*[float(v) for v in lo])
    h = (C.c_float * 3)(*[float(v) for v in hi])
    r = (C.c_int * 4)()
    _check(lib().rt_visible_rect(C.byref(cam), int(width), int(height), l, h, r))
    return tuple(r)


def make_params(width, height, max_level=5, sphere_rays=10, refraction=0.8, sample_mode=0, sample_size=4, exhaustive=False, plane_rays_1d=3, use_bvh=True, glossy_rays=1, texture_debug=False) -> Params:
    p = Params()
    p.width, p.height = int(width), int(height)
    p.max_reflection_level = int(max_level)
    p.sphere_light_ray_count = int(sphere_rays)
    p.glossy_ray_count = int(glossy_rays)
    p.refraction_factor = float(refraction)
    p.sample_mode, p.sample_size = int(sample_mode), int(sample_size)
    p.use_bvh = 1 if use_bvh else 0
    p.exhaustive = 1 if exhaustive else 0
    p.plane_light_ray_count_1d = int(plane_rays_1d)
    p.texture_debug = 1 if texture_debug else 0
    return p


class Context:
    """One rt_ctx (one GPU).  Mirrors the life cycle Scene -> BoundingVolumeHierarchy(&scene) -> renderRayTracing."""

    def __init__(self, device: int = 0):
        self._l = lib()
        self._h = _P()
        _check(self._l.rt_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            self._l.rt_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- scene --
    def upload_scene(self, scene: SceneData, bvh_mode: int = BVH_AUTO):
        pos, nrm = _f32(scene.pos), _f32(scene.nrm)
        ids = np.ascontiguousarray(scene.mesh_id, dtype=np.int32)
        mats = np.ascontiguousarray(scene.mats, dtype=MATERIAL_DTYPE)
        n = pos.shape[0]
        _check(self._l.rt_upload_scene(self._h, pos.ctypes.data if n else None, nrm.ctypes.data if n else None, ids.ctypes.data if n else None, n,
                                       mats.ctypes.data if len(mats) else None, mats.shape[0]))
        _check(self._l.rt_build_bvh(self._h, bvh_mode))
        self.set_lights(scene.point_lights, scene.sphere_lights)
        self.set_spheres(scene.spheres)
        self.set_spot_lights(scene.spot_lights)
        self.set_plane_lights(scene.plane_lights)
        self.set_texcoords(scene.uv)
        self.set_textures(scene.textures, scene.mesh_tex)

    def set_texcoords(self, uv=None):
        if uv is None:
            _check(self._l.rt_set_texcoords(self._h, None))
            return
        uv = _f32(uv).reshape(-1, 6)
        _check(self._l.rt_set_texcoords(self._h, uv.ctypes.data))

    def set_textures(self, textures=None, mesh_tex=None):
        """textures: (H, W, 3) uint8 images, converted like Image::Image (byte / 255.0f, src/image.cpp:57-59)."""
        if not textures:
            _check(self._l.rt_set_textures(self._h, None, 0, None, 0))
            return
        texels = [np.ascontiguousarray(np.asarray(t, np.uint8).astype(np.float32) / np.float32(255.0)) for t in textures]
        arr = (Texture * len(texels))(*[Texture(t.shape[1], t.shape[0], t.ctypes.data) for t in texels])
        mt = np.ascontiguousarray(mesh_tex, np.int32)
        _check(self._l.rt_set_textures(self._h, arr, len(texels), mt.ctypes.data, len(mt)))

    def set_texturing(self, filtering=None, oob_x=OOB_BORDER, oob_y=OOB_BORDER, border=(0.0, 0.0, 0.0)):
        """useTextures and its knobs (src/main.cpp:54-58) for the following frames; filtering=None switches textures off (the
        texture-debug view then samples with the reference's default knobs)."""
        if filtering is None:
            _check(self._l.rt_set_texturing(self._h, None))
            return
        p = TextureParams(int(filtering), int(oob_x), int(oob_y), (C.c_float * 3)(*[float(v) for v in border]))
        _check(self._l.rt_set_texturing(self._h, C.byref(p)))

    def build_bvh(self, mode: int):
        _check(self._l.rt_build_bvh(self._h, mode))

    def bvh_info(self):
        n, d = C.c_int(), C.c_int()
        _check(self._l.rt_bvh_info(self._h, C.byref(n), C.byref(d)))
        return n.value, d.value

    def set_materials(self, mats: np.ndarray):
        mats = np.ascontiguousarray(mats, dtype=MATERIAL_DTYPE)
        _check(self._l.rt_set_materials(self._h, mats.ctypes.data, mats.shape[0]))

    def set_lights(self, point_lights=None, sphere_lights=None):
        pl = _f32(point_lights if point_lights is not None else np.zeros((0, 6))).reshape(-1, 6)
        sl = _f32(sphere_lights if sphere_lights is not None else np.zeros((0, 7))).reshape(-1, 7)
        _check(self._l.rt_set_lights(self._h, pl.ctypes.data if len(pl) else None, len(pl), sl.ctypes.data if len(sl) else None, len(sl)))

    def set_spheres(self, spheres=None):
        """Scene::spheres: (n, 12) rows = centre (3), radius, kd (3), ks (3), shininess, transparency (== rt_sphere)."""
        sp = _f32(spheres if spheres is not None else np.zeros((0, 12))).reshape(-1, 12)
        _check(self._l.rt_set_spheres(self._h, sp.ctypes.data if len(sp) else None, len(sp)))

    def set_spot_lights(self, spot=None):
        sp = _f32(spot if spot is not None else np.zeros((0, 10))).reshape(-1, 10)
        _check(self._l.rt_set_spot_lights(self._h, sp.ctypes.data if len(sp) else None, len(sp)))

    def set_plane_lights(self, plane=None):
        pl = _f32(plane if plane is not None else np.zeros((0, 12))).reshape(-1, 12)
        _check(self._l.rt_set_plane_lights(self._h, pl.ctypes.data if len(pl) else None, len(pl)))

    def set_postprocess(self, post: "PostParams | None"):
        """Post-processing applied on the device at the end of every following frame (None: off)."""
        _check(self._l.rt_set_postprocess(self._h, C.byref(post) if post is not None else None))

    def postprocess(self, rgb, post: "PostParams", via_write_bitmap=False):
        """Screen::postprocessImage (or writeBitmapToFile's bloom + 8-bit conversion) on a host image (H, W, 3)."""
        img = np.array(rgb, np.float32, order="C", copy=True)
        h, w = img.shape[:2]
        rgba = np.zeros((h, w, 4), np.uint8)
        _check(self._l.rt_postprocess(self._h, C.byref(post), img.ctypes.data, w, h, 1 if via_write_bitmap else 0, rgba.ctypes.data if via_write_bitmap else None))
        return (img, rgba) if via_write_bitmap else img

    def postprocess_device(self, post: "PostParams", width: int, height: int, d_rgba=None):
        _check(self._l.rt_postprocess_device(self._h, C.byref(post), d_rgba, int(width), int(height)))

    def set_counters(self, enable: bool):
        _check(self._l.rt_set_counters(self._h, 1 if enable else 0))

    STAGES = ("generate", "extend", "shade", "shadow_point", "shadow_sphere", "resolve", "shadow_plane", "post")

    def set_stage_timing(self, enable: bool):
        _check(self._l.rt_set_stage_timing(self._h, 1 if enable else 0))

    def stage_times(self) -> dict:
        """{stage: (ms, launches)} of the frame completed by the last sync()/render()."""
        ms = (C.c_float * len(self.STAGES))()
        n = (C.c_int * len(self.STAGES))()
        _check(self._l.rt_stage_times(self._h, ms, n))
        return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(self.STAGES)}

    def set_pipeline(self, lanes: int, batches_per_frame: int, min_batch_pixels: int = 1 << 18):
        _check(self._l.rt_set_pipeline(self._h, lanes, batches_per_frame, min_batch_pixels))

    def set_paths(self, mode: int):
        """Bounce levels as whole paths (k_paths): -1 automatic, 0 never, 1 whenever legal."""
        _check(self._l.rt_set_paths(self._h, int(mode)))

    def set_wide(self, mode: int):
        """Eight lanes per ray through the 8-wide tree for the levels >= 1: -1 automatic (small queues), 0 never, 1 always."""
        _check(self._l.rt_set_wide(self._h, int(mode)))

    def measure_fp32_peak(self) -> float:
        """Un-fused FMUL / FADD issue rate of this device in 1e9 lane-instructions per second (microbenchmark)."""
        v = C.c_double(0.0)
        _check(self._l.rt_measure_fp32_peak(self._h, C.byref(v)))
        return float(v.value)

    def set_overlap(self, enable: bool):
        _check(self._l.rt_set_overlap(self._h, 1 if enable else 0))

    def set_batch_rays(self, n: int):
        _check(self._l.rt_set_batch_rays(self._h, int(n)))

    def set_shard(self, rank: int, world: int):
        _check(self._l.rt_set_shard(self._h, rank, world))

    def set_stream(self, cuda_stream: int):
        _check(self._l.rt_set_stream(self._h, _P(cuda_stream)))

    # -- render --
    def render(self, cam: Camera, prm: Params, want_ids: bool = False, rgb_out: np.ndarray | None = None):
        """Host buffers in, host buffers out (the reference-facing call).  Returns (rgb[H,W,3], ids|None, t|None, Stats)."""
        h, w = prm.height, prm.width
        rgb = rgb_out if rgb_out is not None else np.empty((h, w, 3), np.float32)
        ids = np.empty((h, w), np.int32) if want_ids else None
        t = np.empty((h, w), np.float32) if want_ids else None
        st = Stats()
        _check(self._l.rt_render(self._h, C.byref(cam), C.byref(prm), rgb.ctypes.data, ids.ctypes.data if want_ids else None,
                                 t.ctypes.data if want_ids else None, C.byref(st)))
        return rgb, ids, t, st

    def render_host_ptr(self, cam: Camera, prm: Params, rgb_ptr: int) -> Stats:
        """rt_render into caller-owned (ideally pinned) host memory given as an address."""
        st = Stats()
        _check(self._l.rt_render(self._h, C.byref(cam), C.byref(prm), _P(rgb_ptr), None, None, C.byref(st)))
        return st

    def render_shard_host(self, cam: Camera, prm: Params, rgb_mapped_ptr: int) -> Stats:
        """rt_render_shard: this rank's tiles stored straight into the whole image's page-locked, device-mapped host buffer."""
        st = Stats()
        _check(self._l.rt_render_shard(self._h, C.byref(cam), C.byref(prm), _P(rgb_mapped_ptr), C.byref(st)))
        return st

    def render_device(self, cam: Camera, prm: Params, d_rgba: int | None = None):
        _check(self._l.rt_render_device(self._h, C.byref(cam), C.byref(prm), _P(d_rgba) if d_rgba else None))

    def sync(self) -> Stats:
        st = Stats()
        _check(self._l.rt_sync(self._h, C.byref(st)))
        return st

    def framebuffer(self):
        p, w, h = _P(), C.c_int(), C.c_int()
        _check(self._l.rt_framebuffer(self._h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def framebuffer_ipc_handle(self, width: int, height: int) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._l.rt_framebuffer_ipc_handle(self._h, width, height, buf))
        return buf.raw

    def open_peer_framebuffer(self, handle: bytes) -> int:
        p = _P()
        _check(self._l.rt_open_peer_framebuffer(self._h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def set_host_store_rate(self, gbs: float):
        """Device-to-host rate of this rank's link measured while every rank of the job copies at once (GB/s); 0 = built-in assumption."""
        _check(self._l.rt_set_host_store_rate(self._h, float(gbs)))

    def set_gather_target(self, ptr: int):
        """Same-process form of open_peer_framebuffer: `ptr` = the root's framebuffer() pointer right after its export."""
        _check(self._l.rt_set_gather_target(self._h, _P(ptr)))

    def close_peer_framebuffer(self, ptr: int):
        _check(self._l.rt_close_peer_framebuffer(self._h, _P(ptr)))

    def download_rgb_ptr(self, d_rgba: int, width: int, height: int, host_ptr: int):
        """rt_download_rgb into caller-owned (ideally pinned) host memory."""
        _check(self._l.rt_download_rgb(self._h, _P(d_rgba), width, height, _P(host_ptr)))

    def download_rgb(self, d_rgba: int, width: int, height: int) -> np.ndarray:
        rgb = np.empty((height, width, 3), np.float32)
        _check(self._l.rt_download_rgb(self._h, _P(d_rgba), width, height, rgb.ctypes.data))
        return rgb

    def intersect(self, rays: np.ndarray, use_bvh: bool = True):
        """BoundingVolumeHierarchy::intersect for a batch of rays (n, 6) -> (tri_id[n], t[n])."""
        rays = _f32(rays).reshape(-1, 6)
        n = rays.shape[0]
        ids = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        _check(self._l.rt_intersect(self._h, rays.ctypes.data, n, 1 if use_bvh else 0, ids.ctypes.data, t.ctypes.data))
        return ids, t

    def violations(self) -> dict:
        """Checked build (rt_checked_build): indices the kernels found out of range so far, per site; all zero from the default build."""
        counts = (C.c_uint * len(CHECK_SITES))()
        _check(self._l.rt_violations(self._h, counts, len(CHECK_SITES)))
        return dict(zip(CHECK_SITES, (int(v) for v in counts)))

    def violations_selftest(self):
        """Checked build: one deliberate violation at "table_entry" and one at "wide_stack_slot" (a live counter shows them)."""
        _check(self._l.rt_violations_selftest(self._h))


# sites of the checked build's bounds tests, in the order of rt_violations (csrc/rt_types.h: CheckSite)
CHECK_SITES = ("bvh_node", "triangle", "stack_slot", "wide_node", "wide_stack_slot", "accumulator_pixel", "framebuffer_pixel", "texel", "table_entry", "queue_slot")


def checked_build() -> bool:
    return bool(lib().rt_checked_build())


# ---- image sharding (host-side mirror of the device mapping in csrc/rt_kernels.cu: local_to_pixel) ----
TILE_W, TILE_H = 32, 16


def tile_grid(width: int, height: int):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def owner_map(width: int, height: int, world: int) -> np.ndarray:
    """(H, W) array in the Screen layout (row 0 = top): which rank renders each pixel.  Tile g (owner g % world) lies in tile row
    g // tiles_x at column (g % tiles_x + 3 * row) % tiles_x (csrc/rt_types.h: tile_xy), so the tile at (column, row) is
    g = row * tiles_x + (column - 3 * row) % tiles_x."""
    tx, ty = tile_grid(width, height)
    rows, cols = np.arange(ty)[:, None], np.arange(tx)[None, :]
    rot = int(os.environ.get("RTB200_TILE_ROT", "3"))   # developer knob mirrored from the library
    tiles = (rows * tx + (cols - rot * rows) % tx) % world
    full = np.repeat(np.repeat(tiles, TILE_H, 0), TILE_W, 1)[:height, :width]  # indexed [py, px], py = 0 at the bottom
    return full[::-1].copy()                                                     # Screen::setPixel flips y


def local_tile_count(width: int, height: int, rank: int, world: int) -> int:
    tx, ty = tile_grid(width, height)
    total = tx * ty
    return (total - rank + world - 1) // world if total > rank else 0
