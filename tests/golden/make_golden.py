"""Mint the golden fixtures of tests/golden/*.npz.  Run HERE (the container that has /root/reference):

    make -C oracle && make -C raytracer-group27_b200 && python tests/golden/make_golden.py

The reference ships no tests or golden vectors for the render path, so the pins are minted from
the reference's own code: each fixture holds the scene exactly as the OBJ importer produced it from
/root/reference/data (triangle soup, materials), the lights / camera / knobs of the config, and the outputs of
oracle/_ref/libref_oracle.so — the reference's ray_tracing.cpp, bounding_volume_hierarchy.cpp and shadow.cpp
compiled verbatim — for that input: per-pixel colour, closest-hit triangle id and t of the primary ray, and ray
counts.  Scenes whose triangles trip the reference's uninitialised-barycentric bug (teapot, dragon stand-in) take
their colours from the port with defined barycentrics (ids and t still come from the verbatim code).  Every fixture
also records whether port and reference agreed bit for bit when it was minted.
/root/reference does not exist on the GPU box; the tests read only these .npz files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import rtb200  # noqa: E402
from rtb200 import standin  # noqa: E402

DATA = "/root/reference/data/"
OUT = os.path.dirname(os.path.abspath(__file__))

ONLY = sys.argv[1:]
CAM = dict(look_at=(0.0, 0.0, 0.0), euler_deg=(20.0, 20.0, 0.0), dist=3.0, fovy_deg=50.0)  # src/main.cpp:413-414


def mint(name, sc, w, h, *, max_level, sphere_rays=10, sample_mode=0, sample_size=4, colour_from="reference", cam=CAM, plane_rays_1d=3, tex=None, texture_debug=False):
    """tex: dict(filtering, oob_x, oob_y, border) = useTextures on with these knobs (the scene carries uv / textures / mesh_tex);
    texture_debug: renderRayTracing's textureDebugging view (main.cpp:355-356), useTextures off, the knobs of `tex` still apply."""
    if ONLY and not any(k in name for k in ONLY):   # `python make_golden.py tex_mip tex_tri`: mint only the fixtures whose names contain one of the words
        return
    c = rtb200.make_camera(**cam)
    ref, port = oracle.Oracle("reference"), oracle.Oracle("port")
    for o in (ref, port):
        o.set_spheres(sc.spheres)
        o.set_extra_lights(sc.spot_lights, sc.plane_lights, plane_rays_1d)
        if tex:
            o.set_textures(sc.uv, sc.textures, sc.mesh_tex, tex["filtering"], tex["oob_x"], tex["oob_y"], tex["border"], use_textures=not texture_debug)
        else:
            o.set_textures()
    kw = dict(max_level=max_level, sphere_rays=sphere_rays, sample_mode=sample_mode, sample_size=sample_size, use_bvh=True, texture_debug=texture_debug)
    r_rgb, r_ids, r_t, r_st = ref.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, sc.sphere_lights, c, w, h, **kw)
    p_rgb, p_ids, p_t, p_st = port.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, sc.sphere_lights, c, w, h, **kw)
    # the same frame with every BVH search made cull-free: all objects tested in the BVH's own visiting order (what the
    # triangle arithmetic and that order alone define; the reference's AABB test occasionally culls a box whose triangle
    # the triangle test would accept).  ids_x: winners of the primary rays with exact-t ties resolved by that order.
    x_rgb, x_ids, x_t, x_st = port.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, sc.sphere_lights, c, w, h, shadow_exhaustive=True, **kw)
    ids_equal = bool(np.array_equal(r_ids, p_ids))
    t_equal = bool(np.array_equal(r_t.view(np.int32), p_t.view(np.int32)))
    rgb_equal = bool(np.array_equal(r_rgb.view(np.int32), p_rgb.view(np.int32)))
    rgb = r_rgb if colour_from == "reference" else p_rgb
    st = r_st if colour_from == "reference" else p_st
    assert ids_equal and t_equal, name
    assert np.array_equal(x_t.view(np.int32), r_t.view(np.int32)), name
    if colour_from == "reference":
        assert rgb_equal, name
    # exact-t tie census on primary rays: would another triangle give the same t?  (informational)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        pos=sc.pos, nrm=sc.nrm, mesh_id=sc.mesh_id, mats=sc.mats, point_lights=sc.point_lights, sphere_lights=sc.sphere_lights, spheres=sc.spheres,
        spot_lights=sc.spot_lights, plane_lights=sc.plane_lights, plane_rays_1d=plane_rays_1d,
        cam_look_at=np.array(cam["look_at"], np.float32), cam_euler_deg=np.array(cam["euler_deg"], np.float32),
        cam_dist=np.float32(cam["dist"]), cam_fovy_deg=np.float32(cam["fovy_deg"]),
        width=w, height=h, max_level=max_level, sphere_rays=sphere_rays, sample_mode=sample_mode, sample_size=sample_size,
        rgb=rgb, ids=r_ids, t=r_t, ids_x=x_ids,
        primary_rays=st.primary_rays, shadow_queries=st.shadow_queries, secondary_rays=st.secondary_rays,
        rgb_x=x_rgb, primary_rays_x=x_st.primary_rays, shadow_queries_x=x_st.shadow_queries, secondary_rays_x=x_st.secondary_rays,
        colour_from=colour_from, port_equals_reference=np.array([ids_equal, t_equal, rgb_equal]),
        **(dict(uv=sc.uv, mesh_tex=sc.mesh_tex, n_textures=len(sc.textures), tex_filtering=tex["filtering"], tex_oob_x=tex["oob_x"], tex_oob_y=tex["oob_y"],
                tex_border=np.array(tex["border"], np.float32), texture_debug=int(texture_debug), **{f"texture_{k}": t for k, t in enumerate(sc.textures)}) if tex else {}))
    for o in (ref, port):
        o.set_textures()
        o.set_spheres(None)
        o.set_extra_lights(None, None, 3)
    print(f"{name}: {sc.n_tris} tris {w}x{h} rays={st.rays} port==ref ids/t/rgb={ids_equal}/{t_equal}/{rgb_equal} "
          f"rgb_maxdiff={np.abs(r_rgb - p_rgb).max():.3g} hit={np.mean(r_ids >= 0):.3f} | exhaustive shadows: queries {st.shadow_queries}->{x_st.shadow_queries}, "
          f"pixels differing >1e-4: {int((np.abs(x_rgb - rgb).max(axis=2) > 1e-4).sum())}, tie pixels (visiting order != id order): {int((x_ids != r_ids).sum())}")


def with_lights(sc, point=None, sphere=None):
    sc.point_lights = np.array(point if point is not None else np.zeros((0, 6)), np.float32).reshape(-1, 6)
    sc.sphere_lights = np.array(sphere if sphere is not None else np.zeros((0, 7)), np.float32).reshape(-1, 7)
    return sc


def textured_scene(mip_floor=False):
    """mip_floor: the floor gets a 64x64 texture (square power of two: it has a mip pyramid, and the level of detail varies over it)."""
    quad = np.array([[-1.5, -0.6, -1.5, 1.5, -0.6, -1.5, 1.5, -0.6, 1.5], [-1.5, -0.6, -1.5, 1.5, -0.6, 1.5, -1.5, -0.6, 1.5]], np.float32)
    quv = np.array([[-0.5, -0.5, 1.5, -0.5, 1.5, 1.5], [-0.5, -0.5, 1.5, 1.5, -0.5, 1.5]], np.float32)
    cube = rtb200.load_obj(DATA + "cube.obj", False)
    pos = np.concatenate([quad, cube.pos * np.float32(0.4)])
    nrm = np.concatenate([np.tile([0, 1, 0], (2, 3)).astype(np.float32), cube.nrm])
    mesh_id = np.concatenate([np.zeros(2, np.int32), cube.mesh_id + 1]).astype(np.int32)
    mats = np.zeros(1 + len(cube.mats), rtb200.MATERIAL_DTYPE)
    mats["kd"], mats["transparency"] = 0.7, 1
    mats["ks"][0], mats["shininess"][0] = 0.3, 5      # the floor mirrors a little: textured hits of reflection rays
    sc = with_lights(rtb200.SceneData(pos, nrm, mesh_id, mats), point=[[-1, 1, -1, 1, 1, 1]])
    rng = np.random.default_rng(2)
    sc.uv = np.concatenate([quv, rng.random((12, 6)).astype(np.float32)])
    sc.textures = [np.linspace(0, 255, 7 * 5 * 3).reshape(5, 7, 3).astype(np.uint8),
                   ((np.indices((16, 16)).sum(0) % 2)[..., None] * np.array([245, 190, 30]) + 10).astype(np.uint8)]
    sc.mesh_tex = np.array([0, 1, -1, 1, -1, 0, 1], np.int32)
    if mip_floor:
        yy, xx = np.indices((64, 64))
        pat = np.stack([(xx * 4) % 256, (yy * 4 + xx * 2) % 256, ((xx // 8 + yy // 8) % 2) * 200 + 30], -1)
        sc.textures.append((pat ^ rng.integers(0, 64, pat.shape)).astype(np.uint8))
        sc.mesh_tex[0] = 2
    return sc


def main():
    cornell = lambda: rtb200.load_obj(DATA + "CornellBox-Mirror-Rotated.obj", True)
    # C1 at reduced resolution: Cornell, point light (scene.cpp:33), depth 3
    mint("cornell_c1_256", with_lights(cornell(), point=[[0, 0.58, 0, 1, 1, 1]]), 256, 256, max_level=3)
    # C4 at reduced resolution: spherical light (scene.cpp:42), 64 samples, depth 5
    mint("cornell_c4_96", with_lights(cornell(), sphere=[[0, 0.45, 0, 0.1, 1, 1, 1]]), 96, 96, max_level=5, sphere_rays=64)
    # default sphere_light_ray_count (10 -> m=1, n=9), 4-tap AA, non-square frame
    mint("cornell_sph10_aa_80x48", with_lights(cornell(), sphere=[[0, 0.45, 0, 0.1, 1, 1, 1]]), 80, 48, max_level=2, sphere_rays=10, sample_mode=1)
    # multipleRays 16 spp (C5's sampling), ragged frame size (not a multiple of the 32x16 tile)
    mint("cornell_ms16_70x45", with_lights(cornell(), point=[[0, 0.58, 0, 1, 1, 1]]), 70, 45, max_level=2, sample_mode=2, sample_size=16)
    # both light kinds, looking from inside the box
    mint("cornell_inside_128", with_lights(cornell(), point=[[0, 0.3, 0.2, 0.8, 0.7, 0.6]], sphere=[[0.1, 0.45, 0, 0.1, 0.5, 0.5, 1]]), 128, 128,
         max_level=4, sphere_rays=20, cam=dict(look_at=(0.0, 0.0, 0.0), euler_deg=(5.0, 170.0, 0.0), dist=0.9, fovy_deg=70.0))
    # CornellBox preset exactly as loadScene builds it (scene.cpp:28-35): the box + one transparent sphere + point light
    cb = with_lights(cornell(), point=[[0, 0.58, 0, 1, 1, 1]])
    cb.spheres = np.array([[-0.2, 0.15, -0.25, 0.2, 0, 0, 0, 0, 0, 0, 1, 0]], np.float32)
    mint("cornell_preset_sphere_192", cb, 192, 192, max_level=4)
    # Spheres preset (scene.cpp:80-87): three opaque spheres, no triangles, a bright point light; plus a mirror sphere
    sp = rtb200.SceneData(np.zeros((0, 9), np.float32), np.zeros((0, 9), np.float32), np.zeros(0, np.int32), np.zeros(0, rtb200.MATERIAL_DTYPE))
    sp = with_lights(sp, point=[[3, 0, 3, 15, 15, 15]])
    sp.spheres = np.array([[3, -2, 10.2, 1.0, .8, .2, .2, 0, 0, 0, 1, 1], [-2, 2, 4, 2.0, .6, .8, .2, 0, 0, 0, 1, 1], [0, 0, 6, .75, .2, .2, .8, 0, 0, 0, 1, 1],
                           [1.5, 1.0, 5.0, 0.8, .05, .05, .05, .9, .9, .9, 0, 1]], np.float32)
    mint("spheres_preset_160", sp, 160, 160, max_level=3)
    # CornellBoxPlaneLight preset (scene.cpp:44-50): one plane light, 3 x 3 samples; and a denser 4 x 4 variant from inside
    cp = with_lights(cornell())
    cp.plane_lights = np.array([[-0.1, 0.63, -0.1, 0.15, -0.05, 0, 0, 0, 0.2, 1, 1, 1]], np.float32)
    mint("cornell_planelight_160", cp, 160, 160, max_level=3)
    cp2 = with_lights(cornell(), point=[[0.2, 0.2, 0.3, 0.3, 0.3, 0.3]])
    cp2.plane_lights = np.array([[-0.1, 0.63, -0.1, 0.15, -0.05, 0, 0, 0, 0.2, 1, 0.9, 0.8]], np.float32)
    mint("cornell_planelight_inside_96", cp2, 96, 96, max_level=2, plane_rays_1d=4, cam=dict(look_at=(0.0, 0.0, 0.0), euler_deg=(5.0, 170.0, 0.0), dist=0.9, fovy_deg=70.0))
    # Cube preset exactly as loadScene builds it (scene.cpp:18-26): point light + spot light, all faces transparent
    cu = with_lights(rtb200.load_obj(DATA + "cube.obj", False), point=[[-1, 1, -1, 1, 1, 1]])
    cu.spot_lights = np.array([[-1.2, -1, -1, 1, 1.2, 1, 10, 1, 1, 1]], np.float32)
    mint("cube_preset_spot_128", cu, 128, 128, max_level=3)
    # a narrow spot on the monkey: most hits fall outside the cone
    ms = with_lights(rtb200.load_obj(DATA + "monkey-rotated.obj", True))
    ms.spot_lights = np.array([[-1.5, 1.0, -2.0, 1.5, -0.9, 2.0, 12, 1, 0.8, 0.6], [1.0, 1.5, -1.5, -1.0, -1.4, 1.5, 25, 0.2, 0.3, 0.9]], np.float32)
    mint("monkey_spots_128", ms, 128, 128, max_level=2)
    # z-fighting: two coplanar overlapping triangles (exact ties in t over the whole overlap) whose order in the reference's
    # BVH (sorted by centroid y: the green one first) is the reverse of their order in the mesh list (the red one first).
    # The green one is transparent, the red one opaque, so shadow rays (always through the BVH) and primary rays (useBVH)
    # both depend on who wins the tie; a mirror floor adds secondary rays that meet the pair from below
    zpos = np.array([[-1, -0.5, 0, 1, -0.5, 0, 0, 1.5, 0], [-1, -1, 0, 1, -1, 0, 0, 0.8, 0], [-3, -1.2, -3, 0, -1.2, 3, 3, -1.2, -3]], np.float32)
    znrm = np.array([[0, 0, 1] * 3, [0, 0, 1] * 3, [0, 1, 0] * 3], np.float32)
    zm = np.zeros(3, rtb200.MATERIAL_DTYPE)
    zm["kd"] = [[0.8, 0.1, 0.1], [0.1, 0.8, 0.1], [0.3, 0.3, 0.6]]
    zm["ks"] = [[0, 0, 0], [0, 0, 0], [0.5, 0.5, 0.5]]
    zm["shininess"] = [0, 0, 8]
    zm["transparency"] = [1.0, 0.5, 1.0]
    zf = with_lights(rtb200.SceneData(zpos, znrm, np.arange(3, dtype=np.int32), zm), point=[[0.3, 2, 2.5, 1, 1, 1], [0.3, 2, -2.5, 0.7, 0.7, 0.7]])
    mint("zfight_96", zf, 96, 96, max_level=2)
    # student scenes exactly as loadScene builds them (scene.cpp:118-135): many meshes and materials (40 / 84 / 13), specular
    # and one transparent material in AndreasScene; geometry through the restated importer
    for nm, fn, w_, h_, lvl in (("andreas_160x120", "AndreasScene.obj", 160, 120, 2), ("catalin_128x96", "CatalinScene.obj", 128, 96, 1), ("mike_128x96", "MikeScene.obj", 128, 96, 1)):
        mint(nm, with_lights(rtb200.load_obj(DATA + fn, True), point=[[-1, 1, -1, 1, 1, 1]]), w_, h_, max_level=lvl, colour_from="port")
    # textures: a floor quad whose texture coordinates run from -0.5 to 1.5 (every out-of-bounds rule matters) under a cube
    # with random coordinates; a 7x5 gradient and a 16x16 checker; both filters that need no mip level, four rule pairs
    for nm, filt, ox, oy in (("tex_nearest_border_96x80", 0, 0, 0), ("tex_bilinear_clamp_repeat_96x80", 1, 1, 2), ("tex_nearest_repeat_clamp_96x80", 0, 2, 1),
                             ("tex_bilinear_repeat_96x80", 1, 2, 2),
                             # the mip-mapped filters; their level of detail comes from the ray differentials (src/ray_differentials.cpp,
                             # initial state defined in oracle/ref_harness.cpp): the 16x16 checker has a pyramid, the 7x5 gradient has
                             # none and answers white / black
                             ("tex_mipnearest_repeat_96x80", 2, 2, 2), ("tex_mipbilinear_clamp_96x80", 3, 1, 1), ("tex_trilinear_repeat_clamp_96x80", 4, 2, 1)):
        mint(nm, textured_scene(), 96, 80, max_level=2, tex=dict(filtering=filt, oob_x=ox, oob_y=oy, border=(0.2, 0.1, 0.4)))
    # the same with a 64x64 texture on the floor: the level of detail runs over several pyramid levels across the floor, for camera
    # rays (differentials of a default-constructed Ray) and for the rays mirrored by the floor and the cube (differentials from their direction)
    for nm, filt, ox, oy in (("tex_trilinear_floor64_96x80", 4, 2, 2), ("tex_mipnearest_floor64_96x80", 2, 2, 1), ("tex_mipbilinear_floor64_96x80", 3, 1, 2)):
        mint(nm, textured_scene(mip_floor=True), 96, 80, max_level=2, tex=dict(filtering=filt, oob_x=ox, oob_y=oy, border=(0.2, 0.1, 0.4)))
    mint("texdebug_trilinear_floor64_96x80", textured_scene(mip_floor=True), 96, 80, max_level=2, tex=dict(filtering=4, oob_x=2, oob_y=2, border=(0.2, 0.1, 0.4)), texture_debug=True)
    # renderRayTracing(..., textureDebugging = true) (main.cpp:355-356): texel of the corner ray's hit, white without a texture
    mint("texdebug_bilinear_repeat_clamp_96x80", textured_scene(), 96, 80, max_level=2, tex=dict(filtering=1, oob_x=2, oob_y=1, border=(0.2, 0.1, 0.4)), texture_debug=True)
    # Monkey preset: two point lights (scene.cpp:52-57), mirror-ish material
    mint("monkey_192", with_lights(rtb200.load_obj(DATA + "monkey-rotated.obj", True), point=[[-1, 1, -1, 1, 1, 1], [1, -1, -1, 1, 1, 1]]), 192, 192, max_level=3)
    # Cube preset (every material transparent: d 0.452632)
    mint("cube_96", with_lights(rtb200.load_obj(DATA + "cube.obj", False), point=[[-1, 1, -1, 1, 1, 1]]), 96, 96, max_level=3)
    # SingleTriangle preset: 2 triangles, point + magenta spherical light (scene.cpp:11-16)
    tr = rtb200.load_obj(DATA + "tr_def.obj", False)
    tr.mats["kd"] = 1.0
    mint("tr_def_96", with_lights(tr, point=[[-1, 1, -1, 1, 1, 1]], sphere=[[-2.1, 1.24, -0.51, 0.5, 1, 0, 1]]), 96, 96, max_level=3)
    # C2 at reduced resolution: teapot, depth 0; colours from the port (defined barycentrics)
    mint("teapot_c2_256x144", with_lights(rtb200.load_obj(DATA + "teapot.obj", True), point=[[-1, 1, -1, 1, 1, 1]]), 256, 144, max_level=0, colour_from="port")
    # teapot with reflections (ks 0.2)
    mint("teapot_d3_128x72", with_lights(rtb200.load_obj(DATA + "teapot.obj", True), point=[[-1, 1, -1, 1, 1, 1]]), 128, 72, max_level=3, colour_from="port")
    # C3 at reduced resolution on the dragon STAND-IN (data/dragon.obj is absent); geometry is regenerated by
    # rtb200.standin, only the expected outputs are stored
    sc = standin.dragon_standin_scene()
    mint("dragon_standin_c3_160x90", sc, 160, 90, max_level=3, colour_from="port")
    # drop the 6 MB of geometry from that fixture again: the tests regenerate it and check a checksum instead
    path = os.path.join(OUT, "dragon_standin_c3_160x90.npz")
    d = dict(np.load(path))
    if "pos" not in d:   # not minted in this run (name filter)
        return
    d["pos_sum"] = np.float64(d["pos"].astype(np.float64).sum())
    d["pos_sha"] = np.frombuffer(__import__("hashlib").sha256(d["pos"].tobytes()).digest(), np.uint8)
    d["n_tris"] = d["pos"].shape[0]
    del d["pos"], d["nrm"], d["mesh_id"]
    np.savez_compressed(path, **d)


if __name__ == "__main__":
    main()
