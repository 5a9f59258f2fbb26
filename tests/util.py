"""Shared helpers of the test-suite: golden fixture loading and comparison metrics."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_NAMES = ["cornell_planelight_160", "cornell_planelight_inside_96", "cube_preset_spot_128", "monkey_spots_128", "cornell_preset_sphere_192", "spheres_preset_160", "cornell_c1_256", "cornell_c4_96", "cornell_sph10_aa_80x48", "cornell_ms16_70x45", "cornell_inside_128", "monkey_192",
                "cube_96", "zfight_96", "andreas_160x120", "catalin_128x96", "mike_128x96", "tr_def_96", "teapot_c2_256x144", "teapot_d3_128x72", "dragon_standin_c3_160x90"]


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
        if "pos" in d:
            self.scene = rtb200.SceneData(d["pos"], d["nrm"], d["mesh_id"], d["mats"], d["point_lights"], d["sphere_lights"])
            if "spheres" in d:
                self.scene.spheres = d["spheres"]
            if "spot_lights" in d:
                self.scene.spot_lights, self.scene.plane_lights = d["spot_lights"], d["plane_lights"]
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
        return rtb200.make_params(self.w, self.h, self.max_level, self.sphere_rays, 0.8, self.sample_mode, self.sample_size, exhaustive, self.plane_rays_1d, use_bvh)

    def oracle_render(self, kind="port", **kw):
        import oracle
        o = oracle.Oracle(kind)
        s = self.scene
        o.set_spheres(s.spheres)
        o.set_extra_lights(s.spot_lights, s.plane_lights, self.plane_rays_1d)
        return o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, s.sphere_lights, self.camera(), self.w, self.h, max_level=self.max_level,
                        sphere_rays=self.sphere_rays, sample_mode=self.sample_mode, sample_size=self.sample_size, **kw)


def id_mismatch_fraction(a, b):
    return float(np.mean(a != b))


def bits_equal(a, b):
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.int32), np.ascontiguousarray(b).view(np.int32)))
