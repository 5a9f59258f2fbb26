"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes loader for the two CPU checkers declared in oracle_api.h:
  kind "reference": oracle/_ref/libref_oracle.so (reference translation units compiled verbatim + restated harness)
  kind "port"     : oracle/liboracle_port.so     (plain C++ restatement of the whole path)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {"reference": os.path.join(_HERE, "_ref", "libref_oracle.so"), "port": os.path.join(_HERE, "liboracle_port.so")}


class OrcMaterial(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("ks", C.c_float * 3), ("shininess", C.c_float), ("transparency", C.c_float)]


class OrcCamera(C.Structure):
    _fields_ = [("look_at", C.c_float * 3), ("euler", C.c_float * 3), ("dist", C.c_float), ("fovy", C.c_float)]


class OrcParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("max_reflection_level", C.c_int), ("sphere_light_ray_count", C.c_int),
                ("glossy_ray_count", C.c_int), ("refraction_factor", C.c_float), ("use_bvh", C.c_int), ("sample_mode", C.c_int),
                ("sample_size", C.c_int), ("defined_bary", C.c_int), ("x0", C.c_int), ("y0", C.c_int), ("x_step", C.c_int),
                ("y_step", C.c_int), ("num_threads", C.c_int), ("shadow_exhaustive", C.c_int), ("texture_debug", C.c_int)]


class OrcStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_queries", C.c_uint64), ("secondary_rays", C.c_uint64), ("seconds", C.c_double),
                ("threads", C.c_int)]

    @property
    def rays(self) -> int:
        return int(self.primary_rays + self.shadow_queries + self.secondary_rays)


class OrcTexture(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("rgb", C.c_void_p)]


# TextureFiltering (the two without a mip level) and OutOfBoundsRule of src/image.h:18-31
TEX_NEAREST, TEX_BILINEAR, TEX_MIP_NEAREST, TEX_MIP_BILINEAR, TEX_TRILINEAR = range(5)
OOB_BORDER, OOB_CLAMP, OOB_REPEAT = 0, 1, 2


class OrcPost(C.Structure):
    _fields_ = [("filtering_option", C.c_int), ("kernel", C.c_int), ("kernel_repetitions", C.c_int), ("filter_size", C.c_int),
                ("sigma", C.c_float), ("exposure", C.c_float), ("gamma_correction", C.c_int), ("gamma", C.c_float), ("bloom_live", C.c_int)]


# FilteringOption / Kernel of src/screen.h:17-31
FILTER_NONE, FILTER_BLOOM, FILTER_BLOOM_REINHARD, FILTER_BLOOM_EXPOSURE, FILTER_ONLY_LIGHT, FILTER_ONLY_LIGHT_KERNEL = range(6)
KERNEL_BOX, KERNEL_GAUSSIAN = 0, 1

MATERIAL_DTYPE = np.dtype([("kd", np.float32, 3), ("ks", np.float32, 3), ("shininess", np.float32), ("transparency", np.float32)])


def available(kind: str) -> bool:
    return os.path.exists(PATHS[kind])


class Oracle:
    def __init__(self, kind: str = "port"):
        path = PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built: run `make -C oracle`")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.lib.oracle_render.restype = C.c_int
        self.lib.oracle_closest_hit.restype = C.c_int
        self.lib.oracle_kind.restype = C.c_char_p
        self.lib.oracle_set_spheres.restype = None
        self.lib.oracle_set_extra_lights.restype = None
        assert self.lib.oracle_kind().decode() == kind

    def render(self, pos, nrm, mesh_id, mats, point_lights, sphere_lights, cam, width, height, max_level=5, sphere_rays=10,
               refraction=0.8, use_bvh=True, sample_mode=0, sample_size=4, defined_bary=True, want_ids=True, want_rgb=True,
               x0=0, y0=0, x_step=1, y_step=1, num_threads=0, shadow_exhaustive=False, glossy_rays=1, texture_debug=False):
        """cam: dict(look_at, euler (radians), dist, fovy (radians)) or an object with those attributes."""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 9)
        nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 9)
        mesh_id = np.ascontiguousarray(mesh_id, np.int32)
        mats = np.ascontiguousarray(mats, MATERIAL_DTYPE)
        pl = np.ascontiguousarray(point_lights if point_lights is not None else np.zeros((0, 6)), np.float32).reshape(-1, 6)
        sl = np.ascontiguousarray(sphere_lights if sphere_lights is not None else np.zeros((0, 7)), np.float32).reshape(-1, 7)
        oc = OrcCamera()
        get = (lambda k: cam[k]) if isinstance(cam, dict) else (lambda k: getattr(cam, k))
        oc.look_at[:] = [float(v) for v in get("look_at")]
        oc.euler[:] = [float(v) for v in get("euler")]
        oc.dist = float(get("dist"))
        oc.fovy = float(get("fovy"))
        p = OrcParams(width, height, max_level, sphere_rays, int(glossy_rays), refraction, 1 if use_bvh else 0, sample_mode, sample_size,
                      1 if defined_bary else 0, x0, y0, x_step, y_step, num_threads, 1 if shadow_exhaustive else 0, 1 if texture_debug else 0)
        rgb = np.zeros((height, width, 3), np.float32) if want_rgb else None
        ids = np.full((height, width), -1, np.int32) if want_ids else None
        t = np.zeros((height, width), np.float32) if want_ids else None
        st = OrcStats()
        rc = self.lib.oracle_render(C.c_void_p(pos.ctypes.data), C.c_void_p(nrm.ctypes.data), C.c_void_p(mesh_id.ctypes.data), C.c_int(pos.shape[0]),
                                    C.c_void_p(mats.ctypes.data), C.c_int(mats.shape[0]),
                                    C.c_void_p(pl.ctypes.data if len(pl) else None), C.c_int(len(pl)),
                                    C.c_void_p(sl.ctypes.data if len(sl) else None), C.c_int(len(sl)),
                                    C.byref(oc), C.byref(p),
                                    C.c_void_p(rgb.ctypes.data if want_rgb else None), C.c_void_p(ids.ctypes.data if want_ids else None),
                                    C.c_void_p(t.ctypes.data if want_ids else None), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"oracle_render failed with {rc}")
        return rgb, ids, t, st

    def set_spheres(self, spheres):
        """spheres: (n, 12) = centre, radius, kd, ks, shininess, transparency; None / empty clears them."""
        sp = np.ascontiguousarray(spheres if spheres is not None else np.zeros((0, 12)), np.float32).reshape(-1, 12)
        self.lib.oracle_set_spheres(C.c_void_p(sp.ctypes.data if len(sp) else None), C.c_int(len(sp)))

    def set_extra_lights(self, spot=None, plane=None, plane_ray_count_1d=3):
        """spot: (n, 10) = position, direction, angle (degrees), colour; plane: (n, 12) = position, width, height, colour."""
        sp = np.ascontiguousarray(spot if spot is not None else np.zeros((0, 10)), np.float32).reshape(-1, 10)
        pl = np.ascontiguousarray(plane if plane is not None else np.zeros((0, 12)), np.float32).reshape(-1, 12)
        self.lib.oracle_set_extra_lights(C.c_void_p(sp.ctypes.data if len(sp) else None), C.c_int(len(sp)),
                                         C.c_void_p(pl.ctypes.data if len(pl) else None), C.c_int(len(pl)), C.c_int(int(plane_ray_count_1d)))

    def set_textures(self, tri_uv=None, textures=None, mesh_tex=None, filtering=TEX_NEAREST, oob_x=OOB_BORDER, oob_y=OOB_BORDER, border=(0, 0, 0),
                     use_textures=True):
        """Diffuse textures for the following render() calls (None / no textures: off).  tri_uv (n_tris, 6); textures: list of
        (H, W, 3) uint8 arrays, top row first; mesh_tex: texture index per mesh or -1."""
        self.lib.oracle_set_textures.restype = None
        if tri_uv is None or not textures:
            self.lib.oracle_set_textures(None, 0, None, 0, None, 0, 0, 0, 0, 0, None)
            return
        uv = np.ascontiguousarray(tri_uv, np.float32).reshape(-1, 6)
        imgs = [np.ascontiguousarray(t, np.uint8) for t in textures]
        arr = (OrcTexture * len(imgs))(*[OrcTexture(t.shape[1], t.shape[0], t.ctypes.data) for t in imgs])
        mt = np.ascontiguousarray(mesh_tex, np.int32)
        b = np.ascontiguousarray(border, np.float32)
        self.lib.oracle_set_textures(C.c_void_p(uv.ctypes.data), C.c_int(uv.shape[0]), arr, C.c_int(len(imgs)), C.c_void_p(mt.ctypes.data), C.c_int(len(mt)),
                                     C.c_int(1 if use_textures else 0), C.c_int(int(filtering)), C.c_int(int(oob_x)), C.c_int(int(oob_y)), C.c_void_p(b.ctypes.data))

    def postprocess(self, rgb, filtering_option=FILTER_NONE, kernel=KERNEL_BOX, kernel_repetitions=1, filter_size=5, sigma=2.0, exposure=0.5,
                    gamma_correction=False, gamma=2.2, bloom_live=True, via_write_bitmap=False):
        """Screen::postprocessImage (or the bloom + 8-bit conversion of writeBitmapToFile) on an (H, W, 3) float image in the
        Screen layout.  Returns the processed image, plus the RGBA8 rows when via_write_bitmap."""
        img = np.array(rgb, np.float32, order="C", copy=True)
        h, w = img.shape[:2]
        p = OrcPost(int(filtering_option), int(kernel), int(kernel_repetitions), int(filter_size), float(sigma), float(exposure),
                    1 if gamma_correction else 0, float(gamma), 1 if bloom_live else 0)
        rgba = np.zeros((h, w, 4), np.uint8)
        self.lib.oracle_postprocess.restype = C.c_int
        rc = self.lib.oracle_postprocess(C.c_void_p(img.ctypes.data), C.c_int(w), C.c_int(h), C.byref(p), C.c_int(1 if via_write_bitmap else 0),
                                         C.c_void_p(rgba.ctypes.data))
        if rc != 0:
            raise RuntimeError(f"oracle_postprocess failed with {rc}")
        return (img, rgba) if via_write_bitmap else img

    def closest_hit(self, pos, nrm, mesh_id, rays, use_bvh=False):
        """use_bvh: False/0 every object in id order, True/1 the reference-style BVH, 2 (port) every object in that BVH's visiting order."""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 9)
        nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 9)
        mesh_id = np.ascontiguousarray(mesh_id, np.int32)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        ids = np.empty(rays.shape[0], np.int32)
        t = np.empty(rays.shape[0], np.float32)
        rc = self.lib.oracle_closest_hit(C.c_void_p(pos.ctypes.data), C.c_void_p(nrm.ctypes.data), C.c_void_p(mesh_id.ctypes.data), C.c_int(pos.shape[0]),
                                         C.c_void_p(rays.ctypes.data), C.c_int(rays.shape[0]), C.c_int(int(use_bvh)),
                                         C.c_void_p(ids.ctypes.data), C.c_void_p(t.ctypes.data))
        if rc != 0:
            raise RuntimeError(f"oracle_closest_hit failed with {rc}")
        return ids, t
