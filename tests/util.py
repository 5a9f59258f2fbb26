"""Shared helpers of the test-suite: golden fixture loading and comparison metrics."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_NAMES = ["cornell_planelight_160", "cornell_planelight_inside_96", "cube_preset_spot_128", "monkey_spots_128", "cornell_preset_sphere_192", "spheres_preset_160", "cornell_c1_256", "cornell_c4_96", "cornell_sph10_aa_80x48", "cornell_ms16_70x45", "cornell_inside_128", "monkey_192",
                "cube_96", "zfight_96", "tex_nearest_border_96x80", "tex_bilinear_clamp_repeat_96x80", "tex_nearest_repeat_clamp_96x80", "tex_bilinear_repeat_96x80", "tex_mipnearest_repeat_96x80", "tex_mipbilinear_clamp_96x80", "tex_trilinear_repeat_clamp_96x80", "tex_trilinear_floor64_96x80", "tex_mipnearest_floor64_96x80", "tex_mipbilinear_floor64_96x80",
                "texdebug_trilinear_floor64_96x80", "texdebug_bilinear_repeat_clamp_96x80", "andreas_160x120", "catalin_128x96", "mike_128x96", "tr_def_96", "teapot_c2_256x144", "teapot_d3_128x72", "dragon_standin_c3_160x90"]


# Screen settings exercised by the post-processing fixture (tests/golden/make_golden_post.py) and the GPU tests; keyword
# names are those of oracle.Oracle.postprocess.  filtering_option: 0 None, 1 Bloom, 2 +Reinhard, 3 +Exposure, 4 OnlyLight,
# 5 OnlyLightWithKernel; kernel: 0 box, 1 Gaussian.
POST_CONFIGS = [
    dict(filtering_option=1),                                                        # Screen's defaults with the bloom on
    dict(filtering_option=2, kernel=1),                                              # Gaussian, sigma 2, size 5
    dict(filtering_option=3, kernel_repetitions=3, filter_size=2, exposure=0.8),
    dict(filtering_option=2, kernel=1, sigma=0.7, filter_size=3, gamma_correction=True, gamma=1.8),
    dict(filtering_option=4),
    dict(filtering_option=5, kernel=1, kernel_repetitions=4),                        # repetitions are ignored by this option
    dict(filtering_option=1, filter_size=0),                                         # single tap
    dict(filtering_option=1, filter_size=20),                                        # wider than the staged tile path
    dict(filtering_option=2, bloom_live=False, gamma_correction=True),               # gamma only (postprocessImage), bloom only (BMP)
    dict(filtering_option=0, gamma_correction=True, gamma=2.2),
    dict(filtering_option=1, kernel_repetitions=0, sigma=0.0, kernel=1, filter_size=1),  # setter clamps: 1 repetition, sigma 0.001
]
# settings whose arithmetic is +, *, / and comparisons only: bit-exact on the device; the others involve exp / pow
POST_EXACT = [k for k, c in enumerate(POST_CONFIGS) if c["filtering_option"] != 3 and not c.get("gamma_correction")]


class Golden:
    def __init__(self, name):
        import rtb200
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.d = d
        self.w, self.h = int(d["width"]), int(d["height"])
        self.max_level, self.sphere_rays = int(d["max_level"]), int(d["sphere_rays"])
        self.sample_mode, self.sample_size = int(d["sample_mode"]), int(d["sample_size"])
        self.plane_rays_1d = int(d["plane_rays_1d"]) if "plane_rays_1d" in d else 3
        self.rgb, self.ids, self.t = d["rgb"], d["ids"], d["t"]
        # closest-hit ids with exact-t ties resolved by the visiting order of the reference's BVH (what its useBVH=true
        # search returns when no box test culls); self.ids resolves them by global id (its useBVH=false loop)
        self.ids_x = d["ids_x"]
        self.counts = (int(d["primary_rays"]), int(d["shadow_queries"]), int(d["secondary_rays"]))
        # same frame with shadow queries answered exhaustively (oracle_api.h: shadow_exhaustive)
        self.rgb_x = d["rgb_x"]
        self.counts_x = (int(d["primary_rays_x"]), int(d["shadow_queries_x"]), int(d["secondary_rays_x"]))
        self.cam_kw = dict(look_at=tuple(float(v) for v in d["cam_look_at"]), euler_deg=tuple(float(v) for v in d["cam_euler_deg"]),
                           dist=float(d["cam_dist"]), fovy_deg=float(d["cam_fovy_deg"]))
        self.geometry_ok = True
        # useTextures and its knobs (None: textures off, the reference's default)
        self.tex = dict(filtering=int(d["tex_filtering"]), oob_x=int(d["tex_oob_x"]), oob_y=int(d["tex_oob_y"]), border=tuple(float(v) for v in d["tex_border"])) if "tex_filtering" in d else None
        # renderRayTracing's textureDebugging view (main.cpp:355-356): useTextures off, the knobs above still apply
        self.texture_debug = bool(int(d["texture_debug"])) if "texture_debug" in d else False
        if "pos" in d:
            self.scene = rtb200.SceneData(d["pos"], d["nrm"], d["mesh_id"], d["mats"], d["point_lights"], d["sphere_lights"])
            if "spheres" in d:
                self.scene.spheres = d["spheres"]
            if "spot_lights" in d:
                self.scene.spot_lights, self.scene.plane_lights = d["spot_lights"], d["plane_lights"]
            if "uv" in d:
                self.scene.uv, self.scene.mesh_tex = d["uv"], d["mesh_tex"]
                self.scene.textures = [d[f"texture_{k}"] for k in range(int(d["n_textures"]))]
        else:  # dragon stand-in: geometry is regenerated, the fixture only carries a checksum
            from rtb200 import standin
            sc = standin.dragon_standin_scene()
            sc.mats = d["mats"]
            sc.point_lights, sc.sphere_lights = d["point_lights"], d["sphere_lights"]
            self.scene = sc
            self.geometry_ok = hashlib.sha256(np.ascontiguousarray(sc.pos).tobytes()).digest() == bytes(d["pos_sha"])

    def camera(self):
        import rtb200
        return rtb200.make_camera(**self.cam_kw)

    def params(self, exhaustive=False, use_bvh=True):
        import rtb200
        return rtb200.make_params(self.w, self.h, self.max_level, self.sphere_rays, 0.8, self.sample_mode, self.sample_size, exhaustive, self.plane_rays_1d, use_bvh, texture_debug=self.texture_debug)

    def oracle_render(self, kind="port", **kw):
        import oracle
        o = oracle.Oracle(kind)
        s = self.scene
        o.set_spheres(s.spheres)
        o.set_extra_lights(s.spot_lights, s.plane_lights, self.plane_rays_1d)
        if self.tex:
            o.set_textures(s.uv, s.textures, s.mesh_tex, self.tex["filtering"], self.tex["oob_x"], self.tex["oob_y"], self.tex["border"], use_textures=not self.texture_debug)
        else:
            o.set_textures()
        return o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, s.sphere_lights, self.camera(), self.w, self.h, max_level=self.max_level,
                        sphere_rays=self.sphere_rays, sample_mode=self.sample_mode, sample_size=self.sample_size, texture_debug=self.texture_debug, **kw)


def id_mismatch_fraction(a, b):
    return float(np.mean(a != b))


def bits_equal(a, b):
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.int32), np.ascontiguousarray(b).view(np.int32)))


def write_png(path, samples, ctype, depth=8, palette=None, trns=False):
    """Minimal PNG writer for importer tests: `samples` is (H, W, S) integers with S samples per pixel of `depth` bits
    (palette indices for ctype 3); scanlines cycle through the five PNG filter types so the reader's unfiltering is exercised."""
    import struct
    import zlib
    samples = np.asarray(samples)
    h, w, s = samples.shape
    if depth == 16:
        row_bytes = np.stack([samples >> 8, samples & 255], axis=-1).reshape(h, -1).astype(np.uint8)
    elif depth == 8:
        row_bytes = samples.reshape(h, -1).astype(np.uint8)
    else:
        bits = ((samples.reshape(h, -1)[..., None] >> np.arange(depth - 1, -1, -1)) & 1).reshape(h, -1).astype(np.uint8)
        pad = (-bits.shape[1]) % 8
        row_bytes = np.packbits(np.pad(bits, ((0, 0), (0, pad))), axis=1)
    bpp = max(1, s * depth // 8)
    raw = bytearray()
    prev = np.zeros(row_bytes.shape[1], np.int32)
    for y in range(h):
        cur = row_bytes[y].astype(np.int32)
        ft = y % 5
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) // 2
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        raw.append(ft)
        raw += bytes(((cur - pred) & 255).astype(np.uint8))
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", bytes(np.asarray(palette, np.uint8).reshape(-1)))
    if trns:
        out += chunk(b"tRNS", bytes([0, 255]))
    comp = zlib.compress(bytes(raw), 6)
    out += chunk(b"IDAT", comp[: len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(out)
