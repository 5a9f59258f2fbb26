"""GPU parity tests: the CUDA path, called through the C ABI (include/rt_b200.h), against the golden vectors minted
from the reference's own translation units (tests/golden/make_golden.py) and against the CPU oracle run live.

Bars (BASELINE.json north_star): closest-hit triangle ids bit-exact except documented edge/vertex ties (<= 0.01 % of
pixels); t bit-exact on matching ids; per-channel colour within 1e-4; ray counts (primary / shadow queries /
reflection+refraction) identical.
"""
import numpy as np
import pytest

from util import GOLDEN_NAMES, Golden, bits_equal, id_mismatch_fraction

pytestmark = pytest.mark.gpu

COLOUR_TOL = 1e-4      # north_star: per-channel colour within 1e-4
ID_BUDGET = 1e-4       # north_star: <= 0.01 % of pixels may differ (edge / vertex ties)


def _check_against_golden(g, rgb, ids, t, st, what, by_id=False):
    # (1) closest-hit ids and t of the primary rays against the reference's own triangle code: bit-exact.  Equal t resolve
    #     by the visiting order of the reference's BVH (useBVH=true), or by global id (its useBVH=false loop)
    want_ids = g.ids if by_id else g.ids_x
    assert np.array_equal(ids, want_ids), f"{what}: {id_mismatch_fraction(ids, want_ids) * 100:.4f}% of closest-hit ids differ"
    assert bits_equal(t, g.t), f"{what}: t differs"
    # (2) colour and ray counts against the oracle with every BVH search made cull-free — the semantics a
    #     conservative BVH has: every pixel within 1e-4, counts identical
    err_x = np.abs(rgb - g.rgb_x).max(axis=2)
    if by_id:
        err_x[g.ids != g.ids_x] = 0.0 # primary-ray ties that the two visiting orders resolve differently
    assert err_x.max() <= COLOUR_TOL, f"{what}: {(err_x > COLOUR_TOL).sum()} pixels exceed {COLOUR_TOL} vs cull-free oracle (max {err_x.max():.3g})"
    got = (st.primary_rays, st.shadow_queries, st.secondary_rays)
    assert got == g.counts_x, f"{what}: ray counts {got} != {g.counts_x}"
    # (3) colour against the reference as it runs.  Its own 5-level BVH now and then culls a triangle its triangle test
    #     would accept (slab test on flat boxes, axis-parallel directions); the fixture itself shows where: pixels whose
    #     as-run colour differs from the cull-free one.  Everywhere else at most 0.01 % of pixels may differ
    err = np.abs(rgb - g.rgb).max(axis=2)
    err[np.abs(g.rgb - g.rgb_x).max(axis=2) > COLOUR_TOL] = 0.0
    n_bad = int((err > COLOUR_TOL).sum())
    assert n_bad <= max(1, int(ID_BUDGET * err.size)), f"{what}: {n_bad} pixels exceed {COLOUR_TOL} vs the reference (max {err.max():.3g})"


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("bvh", ["ploc", "lbvh", "sah"])
def test_golden(rtb, gpu_ctx, name, bvh):
    g = Golden(name)
    if not g.geometry_ok:
        pytest.skip("stand-in geometry differs on this host's numpy; covered by test_dragon_live_oracle")
    gpu_ctx.upload_scene(g.scene, {"ploc": rtb.BVH_PLOC_DEVICE, "lbvh": rtb.BVH_LBVH_DEVICE, "sah": rtb.BVH_SAH_HOST}[bvh])
    gpu_ctx.set_texturing(**g.tex) if g.tex else gpu_ctx.set_texturing(None)
    try:
        rgb, ids, t, st = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
    finally:
        gpu_ctx.set_texturing(None)
    _check_against_golden(g, rgb, ids, t, st, f"{name}/{bvh}")


@pytest.mark.parametrize("name", ["cornell_c1_256", "cube_96", "tr_def_96", "monkey_192", "cube_preset_spot_128"])
def test_golden_exhaustive(rtb, gpu_ctx, name):
    """Search without the device BVH (loop over every object, rt_params.exhaustive): must give the same frame, in both
    tie orders (the reference's useBVH=true and useBVH=false)."""
    g = Golden(name)
    gpu_ctx.upload_scene(g.scene, rtb.BVH_LBVH_DEVICE)
    for use_bvh in (True, False):
        rgb, ids, t, st = gpu_ctx.render(g.camera(), g.params(exhaustive=True, use_bvh=use_bvh), want_ids=True)
        _check_against_golden(g, rgb, ids, t, st, f"{name}/exhaustive/use_bvh={use_bvh}", by_id=not use_bvh)


def test_tie_order_without_bvh(rtb, gpu_ctx):
    """zfight_96 with useBVH=false: equal t go to the lower global id for camera and reflection rays, while cansee still
    searches through the BVH (shadow.cpp:42), i.e. in its visiting order."""
    g = Golden("zfight_96")
    o_rgb, o_ids, o_t, o_st = g.oracle_render("port", use_bvh=False, shadow_exhaustive=True)
    assert np.array_equal(o_ids, g.ids) and (o_ids != g.ids_x).sum() > 500
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_SAH_HOST, rtb.BVH_LBVH_DEVICE):
        gpu_ctx.upload_scene(g.scene, mode)
        for exhaustive in (False, True):
            rgb, ids, t, st = gpu_ctx.render(g.camera(), g.params(exhaustive=exhaustive, use_bvh=False), want_ids=True)
            assert np.array_equal(ids, o_ids) and bits_equal(t, o_t)
            assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL
            assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (o_st.primary_rays, o_st.shadow_queries, o_st.secondary_rays)


@pytest.mark.parametrize("glossy,sample_mode,size", [(4, 0, (96, 96)), (10, 0, (64, 48)), (3, 2, (40, 30))])
def test_glossy_rays_match_the_port(rtb, gpu_ctx, glossy, sample_mode, size):
    """glossy_ray_count > 1 (main.cpp:204-250).  The reference draws the glossy directions from rand(), shared by its threads, so it
    cannot be the yardstick here; the path defines a counter-based stream instead (rt_b200.h) and the CPU port restates it: ray
    counts equal, colour within 1e-4, on the Cornell box (tall box: ks 0.95, Ns 4; short box dielectric) and with 16 samples
    per pixel (the sample index is part of the path id)."""
    import oracle
    g = Golden("cornell_c1_256")
    w, h = size
    s = g.scene
    o = oracle.Oracle("port")
    o.set_spheres(None)
    o.set_extra_lights(None, None, 3)
    o.set_textures()
    want = o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, None, g.camera(), w, h, max_level=2, sample_mode=sample_mode, sample_size=16,
                    shadow_exhaustive=True, glossy_rays=glossy)
    plain = o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, None, g.camera(), w, h, max_level=2, sample_mode=sample_mode, sample_size=16,
                     shadow_exhaustive=True)
    assert want[3].secondary_rays > 1.2 * plain[3].secondary_rays and np.abs(want[0] - plain[0]).max() > 0.01   # the glossy rays are there
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_SAH_HOST, rtb.BVH_LBVH_DEVICE):
        gpu_ctx.upload_scene(s, mode)
        rgb, ids, t, st = gpu_ctx.render(g.camera(), rtb.make_params(w, h, 2, sample_mode=sample_mode, sample_size=16, glossy_rays=glossy), want_ids=True)
        assert np.array_equal(ids, want[1]) and bits_equal(t, want[2])
        assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (want[3].primary_rays, want[3].shadow_queries, want[3].secondary_rays)
        assert np.abs(rgb - want[0]).max() <= COLOUR_TOL


def test_dragon_live_oracle(rtb, gpu_ctx):
    """Dragon stand-in against the CPU port run on this host (independent of fixture checksums)."""
    import oracle
    from rtb200 import standin
    sc = standin.dragon_standin_scene()
    cam = rtb.make_camera()
    w, h = 96, 54
    o = oracle.Oracle("port")
    o_rgb, o_ids, o_t, o_st = o.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, None, cam, w, h, max_level=3, shadow_exhaustive=True)
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
        gpu_ctx.upload_scene(sc, mode)
        rgb, ids, t, st = gpu_ctx.render(cam, rtb.make_params(w, h, 3), want_ids=True)
        assert np.array_equal(ids, o_ids)
        assert bits_equal(t, o_t)
        assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL
        assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (o_st.primary_rays, o_st.shadow_queries, o_st.secondary_rays)


def test_intersect_matches_oracle(rtb, gpu_ctx):
    """BoundingVolumeHierarchy::intersect for caller rays: random rays into the monkey, BVH and exhaustive."""
    import oracle
    g = Golden("monkey_192")
    rng = np.random.default_rng(5)
    n = 20000
    o = rng.normal(size=(n, 3)).astype(np.float32)
    o = 2.5 * o / np.linalg.norm(o, axis=1, keepdims=True)
    target = rng.uniform(-0.6, 0.6, size=(n, 3)).astype(np.float32)
    d = target - o
    unit = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # Unit directions (|d| = 1 +- ulp, like every ray the renderer makes): BVH and exhaustive search agree with the
    # oracle.  Un-normalised directions: the reference measures t along normalize(d) but evaluates the hit point
    # with d itself (ray_tracing.cpp:65,111), so its answer is only reproducible by exhaustive search.
    for dirs, bvh_modes in ((unit, (True, False)), (d.astype(np.float32), (False,))):
        rays = np.concatenate([o, dirs], 1).astype(np.float32)
        o_ids, o_t = oracle.Oracle("port").closest_hit(g.scene.pos, g.scene.nrm, g.scene.mesh_id, rays, use_bvh=False)
        assert (o_ids >= 0).sum() > 1000
        for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
            gpu_ctx.upload_scene(g.scene, mode)
            for use_bvh in bvh_modes:
                ids, t = gpu_ctx.intersect(rays, use_bvh)
                assert np.array_equal(ids, o_ids), f"mode {mode} bvh {use_bvh}: {(ids != o_ids).sum()} ids differ"
                assert bits_equal(t, o_t)


def test_axis_parallel_and_in_plane_rays(rtb, gpu_ctx):
    """Rays with exactly zero direction components, some of them travelling inside the plane of a flat face (the cube's
    faces and the Cornell walls have zero-thickness boxes): the BVH must still return what exhaustive search returns."""
    import oracle
    for name in ("cube_96", "cornell_c1_256"):
        g = Golden(name)
        lo, hi = g.scene.pos.reshape(-1, 3).min(0), g.scene.pos.reshape(-1, 3).max(0)
        coords = [np.unique(np.concatenate([np.linspace(lo[a] - 0.25, hi[a] + 0.25, 23), g.scene.pos.reshape(-1, 3)[:, a]])).astype(np.float32) for a in range(3)]
        rays = []
        for axis in range(3):
            u, v = (axis + 1) % 3, (axis + 2) % 3
            uu, vv = np.meshgrid(coords[u], coords[v], indexing="ij")
            for sign in (1.0, -1.0):
                o = np.zeros((uu.size, 3), np.float32)
                o[:, u], o[:, v] = uu.ravel(), vv.ravel()
                o[:, axis] = lo[axis] - 1.0 if sign > 0 else hi[axis] + 1.0
                d = np.zeros_like(o)
                d[:, axis] = sign
                rays.append(np.concatenate([o, d], 1))
        rays = np.concatenate(rays, 0).astype(np.float32)
        # many of these rays run exactly through shared edges: equal t on both neighbours, resolved by the reference's
        # visiting order (2: its BVH's order without the box tests; 0: its useBVH=false loop)
        for use_bvh, order in ((True, 2), (False, 0)):
            o_ids, o_t = oracle.Oracle("port").closest_hit(g.scene.pos, g.scene.nrm, g.scene.mesh_id, rays, use_bvh=order)
            assert (o_ids >= 0).sum() > 100
            for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
                gpu_ctx.upload_scene(g.scene, mode)
                ids, t = gpu_ctx.intersect(rays, use_bvh)
                assert np.array_equal(ids, o_ids), f"{name} mode {mode} bvh {use_bvh}: {(ids != o_ids).sum()} of {len(ids)} ids differ"
                assert bits_equal(t, o_t)


def test_far_and_near_cameras(rtb, gpu_ctx):
    """Box padding and prune margins are absolute/relative slacks sized for the reference's camera range (distance 0.1 ..
    100, trackball.cpp:150): the BVH frame must equal the oracle's at both ends."""
    import oracle
    g = Golden("monkey_192")
    o = oracle.Oracle("port")
    s = g.scene
    for dist, fov in ((95.0, 1.5), (0.35, 80.0), (12.0, 10.0)):
        cam = rtb.make_camera(dist=dist, fovy_deg=fov, euler_deg=(35.0, -50.0, 0.0))
        o_rgb, o_ids, o_t, o_st = o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, None, cam, 160, 120, max_level=2, shadow_exhaustive=True)
        assert (o_ids >= 0).mean() > 0.05
        for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
            gpu_ctx.upload_scene(s, mode)
            rgb, ids, t, st = gpu_ctx.render(cam, rtb.make_params(160, 120, 2), want_ids=True)
            assert np.array_equal(ids, o_ids), f"dist {dist}: {(ids != o_ids).sum()} ids differ"
            assert bits_equal(t, o_t)
            assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL
            assert st.rays == o_st.rays


def test_level_of_detail_beyond_the_mip_pyramid(rtb, gpu_ctx):
    """A level of detail past the last level of a texture's pyramid: getBestLevelMipmap rounds it DOWN when that is nearer (to a level
    that does not exist: the two nearest-level filters then answer white, src/image.cpp:268-271, 293-296) and UP to the last level
    otherwise (the 1 x 1 texel); Trilinear answers black (image.cpp:334-341).  Texture coordinates scaled by 150 put every hit of the
    fixture's 16 x 16 and 64 x 64 textures there.  (Found by comparing the port with the reference's own translation units on random
    scenes, tests/test_oracle.py; the port indexed the missing level, the device read past the pyramid.)"""
    import oracle
    g = Golden("tex_mipnearest_floor64_96x80")
    s = g.scene
    uv = (np.asarray(s.uv, np.float32) * np.float32(150.0)).astype(np.float32)
    o = oracle.Oracle("port")
    o.set_spheres(s.spheres)
    o.set_extra_lights(s.spot_lights, s.plane_lights, g.plane_rays_1d)
    gpu_ctx.upload_scene(s, rtb.BVH_PLOC_DEVICE)
    gpu_ctx.set_texcoords(uv)
    try:
        for filtering, debug in ((2, True), (3, True), (2, False), (3, False), (4, False)):
            o.set_textures(uv, s.textures, s.mesh_tex, filtering, 2, 2, (0, 0, 0), use_textures=not debug)
            o_rgb, o_ids, o_t, o_st = o.render(s.pos, s.nrm, s.mesh_id, s.mats, s.point_lights, s.sphere_lights, g.camera(), g.w, g.h, max_level=g.max_level,
                                               sphere_rays=g.sphere_rays, shadow_exhaustive=True, texture_debug=debug)
            if debug:  # the view shows the texel itself: both outcomes occur on the textures that have a pyramid
                hit = o_ids >= 0
                on_pyramid = hit & np.isin(np.asarray(s.mesh_tex)[np.asarray(s.mesh_id)[np.clip(o_ids, 0, None)]], (1, 2))
                white = (o_rgb == 1.0).all(axis=-1)
                assert (white & on_pyramid).sum() > 500 and (~white & on_pyramid).sum() > 500
            gpu_ctx.set_texturing(filtering, 2, 2, (0.0, 0.0, 0.0))
            prm = rtb.make_params(g.w, g.h, g.max_level, g.sphere_rays, 0.8, g.sample_mode, g.sample_size, False, g.plane_rays_1d, True, texture_debug=debug)
            rgb, ids, t, st = gpu_ctx.render(g.camera(), prm, want_ids=True)
            assert np.array_equal(ids, o_ids) and bits_equal(t, o_t), f"filter {filtering}, debug view {debug}"
            assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL, f"filter {filtering}, debug view {debug}: {np.abs(rgb - o_rgb).max()}"
            assert st.rays == o_st.rays
    finally:
        gpu_ctx.set_texturing(None)
        o.set_textures()


def test_random_scenes_match_the_port(rtb, gpu_ctx):
    """Device against the port on the random scenes the port itself is pinned on (tests/test_oracle.py: bit-identical to the reference's own
    translation units there): random triangle soups with shiny and transparent materials, one or two point lights, sometimes a spherical
    light, a random camera, depth 3."""
    import oracle
    from test_oracle import _random_scene
    o = oracle.Oracle("port")
    o.set_spheres(None)
    o.set_extra_lights(None, None, 3)
    o.set_textures()
    for seed in range(1000, 1012):
        (pos, nrm, mesh, mats, pl, sl, c), rng = _random_scene(seed)
        cam = rtb.make_camera(look_at=c["look_at"], euler_deg=tuple(float(v) for v in rng.uniform(-60, 60, 3)), dist=c["dist"], fovy_deg=float(rng.uniform(35, 65)))
        sc = rtb.SceneData(pos, nrm, mesh, mats, pl, sl if sl is not None else np.zeros((0, 7), np.float32))
        o_rgb, o_ids, o_t, o_st = o.render(pos, nrm, mesh, mats, pl, sl, cam, 64, 48, max_level=3, sphere_rays=6, shadow_exhaustive=True)
        assert (o_ids >= 0).mean() > 0.02
        gpu_ctx.upload_scene(sc, rtb.BVH_PLOC_DEVICE)
        rgb, ids, t, st = gpu_ctx.render(cam, rtb.make_params(64, 48, 3, 6), want_ids=True)
        assert np.array_equal(ids, o_ids) and bits_equal(t, o_t), f"seed {seed}: {(ids != o_ids).sum()} ids differ"
        assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL, f"seed {seed}: colour differs by {np.abs(rgb - o_rgb).max()}"
        assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (o_st.primary_rays, o_st.shadow_queries, o_st.secondary_rays), f"seed {seed}"


def test_full_size_bvh_equals_exhaustive(rtb, gpu_ctx):
    """Size-independent property at C1's full size (1024x1024, depth 3): the BVH frame equals the exhaustive frame
    bit for bit (ids, t) and to round-off in colour — both use the reference's triangle arithmetic."""
    g = Golden("cornell_c1_256")
    gpu_ctx.upload_scene(g.scene, rtb.BVH_LBVH_DEVICE)
    cam = g.camera()
    a = gpu_ctx.render(cam, rtb.make_params(1024, 1024, 3), want_ids=True)
    b = gpu_ctx.render(cam, rtb.make_params(1024, 1024, 3, exhaustive=True), want_ids=True)
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])
    assert np.abs(a[0] - b[0]).max() <= 1e-6
    assert a[3].rays == b[3].rays
    # downsampled consistency with the 256x256 golden: every 4th pixel corner coincides with a golden pixel corner
    # (pixel x of the 1024 frame has ndc x/1024*2-1 == (x/4)/256*2-1 for x % 4 == 0)
    sub_ids = a[1][3::4, 0::4]  # rows are flipped: row (H-1-y); y % 4 == 0 <=> row index % 4 == 3
    assert np.array_equal(sub_ids, g.ids)


def test_batched_frame_equals_single_batch(rtb, gpu_ctx):
    """Wavefront batching bounds ray-state memory; the image must not depend on the batch size."""
    g = Golden("cornell_inside_128")
    gpu_ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
    a = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
    gpu_ctx.set_batch_rays(2048)
    try:
        b = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
    finally:
        gpu_ctx.set_batch_rays(1 << 24)
    assert b[3].batches > 1
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])
    assert np.abs(a[0] - b[0]).max() <= 1e-6
    assert a[3].rays == b[3].rays


def test_sharded_tiles_compose(rtb, gpu_ctx):
    """Interleaved-tile sharding: ranks 0..3 of a world of 4 rendered one after the other into the same framebuffer
    give the single-GPU frame (multi-rank emulated on one GPU, sequentially)."""
    g = Golden("cornell_c1_256")
    gpu_ctx.upload_scene(g.scene, rtb.BVH_LBVH_DEVICE)
    cam, prm = g.camera(), g.params()
    full = gpu_ctx.render(cam, prm)[0]
    acc = np.zeros_like(full)
    rays = 0
    try:
        for r in range(4):
            gpu_ctx.set_shard(r, 4)
            gpu_ctx.render_device(cam, prm)
            st = gpu_ctx.sync()
            rays += st.rays
        ptr, w, h = gpu_ctx.framebuffer()
        acc = gpu_ctx.download_rgb(ptr, w, h)
    finally:
        gpu_ctx.set_shard(0, 1)
    assert np.abs(acc - full).max() <= 1e-6
    assert rays == sum(g.counts_x)


def test_gather_frames_store_only_rows_with_hits(rtb, gpu_ctx):
    """The gather of a sharded frame (rt_b200.h, "The gather"): a root and a peer context — two ranks of a world of 2, here on one
    GPU — render gather frames into the root's exported framebuffer; the peer stores only the tile rows that hold a hit, the root's
    background fill of the alternate half supplies the rest.  Every gathered frame equals the single-context frame, also when the
    camera changes between frames (rows that had hits one frame and have none the next)."""
    g = Golden("monkey_192")
    cam_a, prm = g.camera(), rtb.make_params(384, 256, 2)
    cam_b = rtb.make_camera(euler_deg=(20.0, 65.0, 0.0), dist=4.5)
    gpu_ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
    want = {}
    for name, cam in (("a", cam_a), ("b", cam_b)):
        gpu_ctx.render_device(cam, prm)
        gpu_ctx.sync()
        ptr, w, h = gpu_ctx.framebuffer()
        want[name] = gpu_ctx.download_rgb(ptr, w, h)
    assert np.abs(want["a"] - want["b"]).max() > 0.1
    peer = rtb.Context(0)
    try:
        peer.upload_scene(g.scene, rtb.BVH_SAH_HOST)
        gpu_ctx.set_shard(0, 2)
        peer.set_shard(1, 2)
        gpu_ctx.framebuffer_ipc_handle(prm.width, prm.height)
        base = gpu_ctx.framebuffer()[0]
        peer.set_gather_target(base)
        for k, name in enumerate("abbaab"):
            cam = cam_a if name == "a" else cam_b
            gpu_ctx.render_device(cam, prm)
            peer.render_device(cam, prm, base)
            st_root, st_peer = gpu_ctx.sync(), peer.sync()   # (sync on every rank + barrier between frames in a real job)
            ptr, w, h = gpu_ctx.framebuffer()
            got = gpu_ctx.download_rgb(ptr, w, h)
            assert np.abs(got - want[name]).max() <= 1e-6, f"frame {k} ({name})"
            n_local = (prm.width * prm.height) // 2
            assert st_root.gather_bytes == 0 and 0 < st_peer.gather_bytes < n_local * 16 // 2   # well under half of the peer's pixels travel
            assert st_root.rays + st_peer.rays > prm.width * prm.height
    finally:
        peer.close()
        gpu_ctx.set_shard(0, 1)
    # a plain frame on the (still exported) root does not go through the gather halves
    gpu_ctx.render_device(cam_a, prm)
    gpu_ctx.sync()
    ptr, w, h = gpu_ctx.framebuffer()
    assert np.abs(gpu_ctx.download_rgb(ptr, w, h) - want["a"]).max() <= 1e-6


@pytest.mark.parametrize("name", ["monkey_192", "monkey_spots_128", "teapot_d3_128x72", "spheres_preset_160", "cornell_c1_256"])
def test_paths_and_levels_give_the_same_frame(rtb, gpu_ctx, name):
    """The bounce levels traced per level (extend / shade / shadow kernels) and as whole paths (k_paths, rt_set_paths) are the same
    computation: equal ray counts, colours to summation order.  Scenes with a transparent material or other light kinds are not legal for
    paths and must silently take the per-level kernels (cornell_c1_256: the short box is a dielectric)."""
    g = Golden(name)
    gpu_ctx.upload_scene(g.scene)
    out = []
    try:
        for mode in (0, 1):
            gpu_ctx.set_paths(mode)
            rgb, ids, t, st = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
            _check_against_golden(g, rgb, ids, t, st, f"{name}/paths={mode}")
            out.append((rgb, st))
    finally:
        gpu_ctx.set_paths(-1)
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-6
    assert out[0][1].rays == out[1][1].rays
    legal = name != "cornell_c1_256"
    assert (out[1][1].kernel_launches < out[0][1].kernel_launches) == legal   # fewer launches exactly when the path kernel ran


@pytest.mark.parametrize("name", ["monkey_192", "monkey_spots_128", "teapot_d3_128x72", "cornell_c1_256", "cornell_preset_sphere_192", "zfight_96", "dragon_standin_c3_160x90",
                                  "cornell_inside_128"])
def test_eight_lanes_per_ray_give_the_same_frame(rtb, gpu_ctx, name):
    """The bounce levels traced with one lane per ray through the binary tree and with eight lanes per ray through the 8-wide tree
    (rt_set_wide; k_extend_wide, k_shadow_point_wide) find the same hits bit for bit — same triangle arithmetic, same tie rule (zfight_96:
    coplanar triangles) — so ray counts are equal and colours agree to summation order; both against the golden frame.  Forced on for every
    level >= 1 here; on its own the library picks it per level for queues that were small in the previous frame."""
    g = Golden(name)
    if not g.geometry_ok:
        pytest.skip("stand-in geometry differs on this host's numpy")
    gpu_ctx.upload_scene(g.scene)
    out = []
    try:
        gpu_ctx.set_paths(0)
        for mode in (0, 1):
            gpu_ctx.set_wide(mode)
            rgb, ids, t, st = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
            _check_against_golden(g, rgb, ids, t, st, f"{name}/wide={mode}")
            out.append((rgb, st))
    finally:
        gpu_ctx.set_wide(-1)
        gpu_ctx.set_paths(-1)
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-6
    assert (out[0][1].primary_rays, out[0][1].shadow_queries, out[0][1].secondary_rays) == (out[1][1].primary_rays, out[1][1].shadow_queries, out[1][1].secondary_rays)


def test_wide_tree_answers_like_the_binary_one(rtb, gpu_ctx):
    """rt_intersect through the 8-wide tree (use_bvh = 2, eight lanes per ray) against the binary tree on incoherent rays that start on the
    surface: ids and t bit for bit."""
    from rtb200 import standin
    sc = standin.dragon_standin_scene()
    gpu_ctx.upload_scene(sc)
    rng = np.random.default_rng(11)
    n = 50000
    tri = rng.integers(0, sc.n_tris, n)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    p = (sc.pos.reshape(-1, 3, 3)[tri] * w[..., None]).sum(1)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([p + 0.01 * d, d], 1).astype(np.float32)
    ids2, t2 = gpu_ctx.intersect(rays, use_bvh=True)
    ids8, t8 = np.empty(n, np.int32), np.empty(n, np.float32)
    assert rtb.lib().rt_intersect(gpu_ctx._h, rays.ctypes.data, n, 2, ids8.ctypes.data, t8.ctypes.data) == 0
    assert np.array_equal(ids2, ids8) and bits_equal(t2, t8) and (ids2 >= 0).mean() > 0.3


def test_queue_overflow_is_clean_and_retried(rtb, gpu_ctx):
    """A wavefront queue that overflows (views dominated by dielectrics or glossy surfaces can outgrow any fixed head-room) must end
    the frame with RT_ERR_OVERFLOW and nothing else: consumers clamp their item counts to the queues' capacities and the rest of
    the batch is skipped.  The device-resident call reports it; rt_render renders again with twice the head-room per attempt and
    returns the frame.  RTB200_QUEUE_SCALE (read at context creation) shrinks the head-room to provoke it."""
    import os
    g = Golden("cornell_c1_256")
    cam, prm = g.camera(), g.params()
    gpu_ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
    want = gpu_ctx.render(cam, prm, want_ids=True)
    os.environ["RTB200_QUEUE_SCALE"] = "0.05"
    try:
        small = rtb.Context(0)
    finally:
        del os.environ["RTB200_QUEUE_SCALE"]
    try:
        small.upload_scene(g.scene, rtb.BVH_SAH_HOST)
        for _ in range(2):   # twice: the context stays usable
            small.render_device(cam, prm)
            with pytest.raises(rtb.RtError, match="overflow"):
                small.sync()
        rgb, ids, t, st = small.render(cam, prm, want_ids=True)
        assert np.array_equal(ids, want[1]) and bits_equal(t, want[2])
        assert np.abs(rgb - want[0]).max() <= 1e-6
        assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (want[3].primary_rays, want[3].shadow_queries, want[3].secondary_rays)
        # glossy_ray_count at the reference's default of 10 (main.cpp:126) renders as well: the queues are sized from it
        glossy = small.render(cam, rtb.make_params(96, 96, 3, glossy_rays=10))
        assert glossy[3].secondary_rays > want[3].secondary_rays * (96 * 96) / (g.w * g.h)
    finally:
        small.close()


def test_errors_are_loud(rtb, gpu_ctx):
    g = Golden("tr_def_96")
    gpu_ctx.upload_scene(g.scene)
    for count in (0, 41):   # the reference's slider goes from 1 to 40 (main.cpp:530)
        bad = g.params()
        bad.glossy_ray_count = count
        with pytest.raises(rtb.RtError):
            gpu_ctx.render(g.camera(), bad)
    with pytest.raises(rtb.RtError):
        gpu_ctx.render(g.camera(), rtb.make_params(0, 10))
    bad_ids = rtb.SceneData(g.scene.pos, g.scene.nrm, g.scene.mesh_id + 5, g.scene.mats)   # mesh ids outside the material table
    with pytest.raises(rtb.RtError):
        rtb.Context(0).upload_scene(bad_ids)
    with pytest.raises(rtb.RtError):
        gpu_ctx.set_spheres(np.zeros((65, 12), np.float32))                                  # more than 64 spheres
    # a scene without any primitive is legal (the reference renders it black) and must not crash
    empty = rtb.SceneData(np.zeros((0, 9), np.float32), np.zeros((0, 9), np.float32), np.zeros(0, np.int32), np.zeros(0, rtb.MATERIAL_DTYPE))
    gpu_ctx.upload_scene(empty)
    rgb, ids, t, st = gpu_ctx.render(g.camera(), rtb.make_params(64, 48, 2), want_ids=True)
    assert rgb.max() == 0 and (ids == -1).all() and st.rays == 64 * 48


def test_full_size_two_bvhs_agree_c3(rtb, gpu_ctx):
    """C3 at BASELINE size (3840x2160, depth 3) on the dragon stand-in: three unrelated hierarchies (device PLOC + SAH, device LBVH, host SAH)
    must produce the same frame — closest-hit ids and t bit for bit, identical ray counts, colour to accumulation
    round-off.  Any box test that culled a triangle the exact test accepts would show up as a difference."""
    from rtb200 import standin
    sc = standin.dragon_standin_scene()
    cam, prm = rtb.make_camera(), rtb.make_params(3840, 2160, 3)
    out = []
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
        gpu_ctx.upload_scene(sc, mode)
        out.append(gpu_ctx.render(cam, prm, want_ids=True))
    a_rgb, a_ids, a_t, a_st = out[0]
    for b_rgb, b_ids, b_t, b_st in out[1:]:
        assert np.array_equal(a_ids, b_ids) and bits_equal(a_t, b_t)
        assert (a_st.primary_rays, a_st.shadow_queries, a_st.secondary_rays) == (b_st.primary_rays, b_st.shadow_queries, b_st.secondary_rays)
        assert np.abs(a_rgb - b_rgb).max() <= 1e-6
    assert a_st.primary_rays == 3840 * 2160
    assert (a_ids >= 0).mean() > 0.1
    # strided agreement with the golden minted at 160x90 (every 24th pixel corner coincides)
    g = Golden("dragon_standin_c3_160x90")
    if g.geometry_ok:
        assert np.array_equal(a_ids[23::24, 0::24], g.ids)
        assert bits_equal(a_t[23::24, 0::24], g.t)
        assert np.abs(a_rgb[23::24, 0::24] - g.rgb).max() <= COLOUR_TOL


def test_full_size_teapot_c2_bvh_equals_exhaustive(rtb, gpu_ctx):
    """C2 at BASELINE size (teapot 1920x1080, Phong + hard shadows): BVH frame == exhaustive frame."""
    g = Golden("teapot_c2_256x144")
    gpu_ctx.upload_scene(g.scene, rtb.BVH_LBVH_DEVICE)
    cam = g.camera()
    a = gpu_ctx.render(cam, rtb.make_params(1920, 1080, 0), want_ids=True)
    b = gpu_ctx.render(cam, rtb.make_params(1920, 1080, 0, exhaustive=True), want_ids=True)
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])
    assert np.abs(a[0] - b[0]).max() <= 1e-6 and a[3].rays == b[3].rays
    # against the reference: pixel corner (15k, 15m) of the 1920x1080 frame is corner (2k, 2m) of the 256x144 golden (x / W and the
    # aspect ratio are the same floats), so those pixels must carry the golden's ids, t and colour
    rgb, ids, t = a[0][::-1][0::15, 0::15], a[1][::-1][0::15, 0::15], a[2][::-1][0::15, 0::15]   # [::-1]: row 0 = pixel row y = 0
    assert ids.shape == (72, 128) and (ids >= 0).mean() > 0.05
    assert np.array_equal(ids, g.ids_x[::-1][0::2, 0::2]) and bits_equal(t, g.t[::-1][0::2, 0::2])
    assert np.abs(rgb - g.rgb_x[::-1][0::2, 0::2]).max() <= COLOUR_TOL


def test_c5_lattice_small(rtb, gpu_ctx):
    """BASELINE.json configs[4] in small: a 2x2x2 lattice of a reduced stand-in (8 x 2 160 triangles, built by the same code as the
    8x8x8 lattice of 86 880-triangle copies), multipleRays with 16 samples per pixel, depth 3, against the CPU port: closest-hit
    ids and t bit-exact, ray counts equal, colour within 1e-4; both builders."""
    import oracle
    from rtb200 import standin
    sc = standin.dragon_lattice_scene(2, nu=72, nv=15)
    assert sc.n_tris == 8 * 2 * 72 * 15
    cam, (w, h) = rtb.make_camera(), (112, 64)
    o = oracle.Oracle("port")
    o.set_spheres(None)
    o.set_extra_lights(None, None, 3)
    o.set_textures()
    o_rgb, o_ids, o_t, o_st = o.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, None, cam, w, h, max_level=3, sample_mode=2, sample_size=16,
                                       shadow_exhaustive=True)
    assert o_st.primary_rays == w * h * 16 and (o_ids >= 0).mean() > 0.03 and o_st.secondary_rays > 1000
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
        gpu_ctx.upload_scene(sc, mode)
        rgb, ids, t, st = gpu_ctx.render(cam, rtb.make_params(w, h, 3, sample_mode=2, sample_size=16), want_ids=True)
        assert np.array_equal(ids, o_ids) and bits_equal(t, o_t)
        assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == (o_st.primary_rays, o_st.shadow_queries, o_st.secondary_rays)
        assert np.abs(rgb - o_rgb).max() <= COLOUR_TOL


def test_c5_lattice_full_mesh_subset_equals_exhaustive(rtb, gpu_ctx):
    """C5's own mesh — the 8x8x8 lattice, 44.5 M triangles, the default device builder (PLOC + SAH, rt_ploc.cu) — at a size the exhaustive search can afford: a 32x18 frame with
    16 samples per pixel, depth 1.  The BVH frame must equal the exhaustive frame (every object tested for every ray) in ids, t, ray
    counts and colour; a box of the 44.5 M-triangle tree that culled a triangle the exact test accepts would show up here."""
    from rtb200 import standin
    sc = standin.dragon_lattice_scene(8)
    assert sc.n_tris == 512 * 86880
    gpu_ctx.upload_scene(sc, rtb.BVH_PLOC_DEVICE)
    del sc
    cam = rtb.make_camera()
    a = gpu_ctx.render(cam, rtb.make_params(32, 18, 1, sample_mode=2, sample_size=16), want_ids=True)
    b = gpu_ctx.render(cam, rtb.make_params(32, 18, 1, sample_mode=2, sample_size=16, exhaustive=True), want_ids=True)
    assert (a[1] >= 0).mean() > 0.05 and a[3].secondary_rays > 100
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])
    assert (a[3].primary_rays, a[3].shadow_queries, a[3].secondary_rays) == (b[3].primary_rays, b[3].shadow_queries, b[3].secondary_rays)
    assert np.abs(a[0] - b[0]).max() <= 1e-6
    gpu_ctx.upload_scene(Golden("cube_96").scene)   # give the 10 GB back


def test_full_size_c4_soft_shadow_sums(rtb, gpu_ctx):
    """C4 at BASELINE size (Cornell 2048x2048, spherical light, 64 samples, depth 5): the frame does not depend on the BVH
    and every shadow sample is accounted for (64 queries per hit plus the extra iterations behind the glass box)."""
    g = Golden("cornell_c4_96")
    cam, prm = g.camera(), rtb.make_params(2048, 2048, 5, sphere_rays=64)
    out = []
    for mode in (rtb.BVH_PLOC_DEVICE, rtb.BVH_LBVH_DEVICE, rtb.BVH_SAH_HOST):
        gpu_ctx.upload_scene(g.scene, mode)
        out.append(gpu_ctx.render(cam, prm, want_ids=True))
    a = out[0]
    for b in out[1:]:
        assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])
        assert np.abs(a[0] - b[0]).max() <= 2e-6
        assert a[3].rays == b[3].rays
    hits = int((a[1] >= 0).sum()) + int(a[3].secondary_rays)  # an upper bound of shaded hits: primary hits + all children
    assert a[3].shadow_queries >= 64 * int((a[1] >= 0).sum())
    assert a[3].shadow_queries <= 64 * hits * 3
    assert np.isfinite(a[0]).all() and a[0].max() <= 2.0
    # against the reference: pixel corner (64j, 64m) of the 2048x2048 frame is corner (3j, 3m) of the 96x96 golden
    rgb, ids, t = a[0][::-1][0::64, 0::64], a[1][::-1][0::64, 0::64], a[2][::-1][0::64, 0::64]
    assert ids.shape == (32, 32)
    assert np.array_equal(ids, g.ids_x[::-1][0::3, 0::3]) and bits_equal(t, g.t[::-1][0::3, 0::3])
    assert np.abs(rgb - g.rgb_x[::-1][0::3, 0::3]).max() <= COLOUR_TOL


CPP_RENDER = r"""
// The reference's main.cpp flow (src/main.cpp:410-420, 513-521) against the drop-in headers.
#include "bounding_volume_hierarchy.h"
#include "render.h"
#include "screen.h"
#include "trackball.h"
#include <cstdio>
#include <fstream>
int main(int argc, char** argv) {
    const glm::ivec2 windowResolution{256, 256};
    Window window{"Final Project - Part 2", windowResolution, OpenGLVersion::GL2};
    Screen screen{windowResolution};
    Trackball camera{&window, glm::radians(50.0f), 3.0f};
    camera.setCamera(glm::vec3(0.0f, 0.0f, 0.0f), glm::radians(glm::vec3(20.0f, 20.0f, 0.0f)), 3.0f);
    Scene scene = loadScene(Custom, argv[1]);          // custom.obj + point light (-1,1,-1) (scene.cpp:88-98)
    BoundingVolumeHierarchy bvh{&scene};
    max_reflection_level = 3;
    glossy_ray_count = 1;
    renderRayTracing(scene, camera, bvh, screen);
    std::ofstream f(argv[2], std::ios::binary);
    f.write(reinterpret_cast<const char*>(screen.pixels().data()), sizeof(glm::vec3) * screen.pixels().size());
    Ray r = camera.generateRay(glm::vec2(0.0f, 0.0f));
    HitInfo hi;
    const bool hit = bvh.intersect(r, hi, true);
    std::printf("%d %d %.9g %llu\n", hit ? 1 : 0, hi.triangle_index, r.t, lastRenderTimings.rays);
    // second frame: a light bright enough to bloom, Screen configured through the reference's setters (main.cpp:560-640)
    scene.pointLights[0].color = glm::vec3(6.0f);
    screen.setBloomFilter(FilteringOption::BloomWithReinhardHdr);
    screen.setKernel(Kernel::GaussianKernel);
    screen.setFilterSize(3);
    screen.setSigma(1.5f);
    screen.setBloomFilterLive(true);
    renderRayTracing(scene, camera, bvh, screen);    // ends with the post-processing (main.cpp:397-398)
    std::ofstream f2(argv[3], std::ios::binary);
    f2.write(reinterpret_cast<const char*>(screen.pixels().data()), sizeof(glm::vec3) * screen.pixels().size());
    screen.writeBitmapToFile(argv[4]);                // blooms the pixels once more, then writes 8-bit BGRA (screen.cpp:40-53)
    // third frame: textures on (the "Use Textures" panel of main.cpp:598-610), bilinear, repeat in x, clamp in y
    screen.setBloomFilter(FilteringOption::None);
    scene.pointLights[0].color = glm::vec3(1.0f);
    useTextures = true;
    textureFiltering = TextureFiltering::Bilinear;
    outOfBoundsRuleX = OutOfBoundsRule::Repeat;
    outOfBoundsRuleY = OutOfBoundsRule::Clamp;
    renderRayTracing(scene, camera, bvh, screen);
    std::ofstream f3(argv[5], std::ios::binary);
    f3.write(reinterpret_cast<const char*>(screen.pixels().data()), sizeof(glm::vec3) * screen.pixels().size());
    // fourth frame: the texture-debug view (main.cpp:753-757), useTextures off again, the same knobs
    useTextures = false;
    renderRayTracing(scene, camera, bvh, screen, true);
    std::ofstream f4(argv[6], std::ios::binary);
    f4.write(reinterpret_cast<const char*>(screen.pixels().data()), sizeof(glm::vec3) * screen.pixels().size());
    return 0;
}
"""


def test_cpp_drop_in_renders_the_same_frame(rtb, gpu_ctx, tmp_path):
    """renderRayTracing / BoundingVolumeHierarchy / Screen / Trackball of host/ (C++), built with g++ against the library,
    give bit for bit the frame the C ABI gives through Python for the same OBJ, camera and knobs."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host, lib = os.path.join(root, "raytracer-group27_b200", "host"), os.path.join(root, "raytracer-group27_b200")
    # a small closed mesh with a mirror-ish and a matte material
    obj = ["mtllib custom.mtl", "o box"]
    v = [(x, y, z) for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)]
    obj += [f"v {a} {b} {c}" for a, b, c in v]
    obj += ["vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1", "vt 2.5 -0.5", "vt -1 2"]
    faces = [(1, 2, 4, 3), (5, 7, 8, 6), (1, 5, 6, 2), (3, 4, 8, 7), (1, 3, 7, 5), (2, 6, 8, 4)]
    for i, f in enumerate(faces):
        obj.append(f"usemtl m{i % 2}")
        obj.append("f " + " ".join(f"{k}/{t}" for k, t in zip(f, (1, 2, 3, 4))))
    obj += ["o floor", "v -2 -0.6 -2", "v 2 -0.6 -2", "v 2 -0.6 2", "v -2 -0.6 2", "usemtl m0", "f 9/5 10/2 11/6 12/4"]
    (tmp_path / "custom.obj").write_text("\n".join(obj) + "\n")
    (tmp_path / "custom.mtl").write_text("newmtl m0\nKd 0.7 0.6 0.5\nKs 0.4 0.4 0.4\nNs 20\nmap_Kd tiles.png\nnewmtl m1\nKd 0.2 0.5 0.8\nKs 0 0 0\nNs 5\n")
    from util import write_png
    write_png(tmp_path / "tiles.png", np.random.default_rng(8).integers(0, 256, (8, 8, 3)), 2)
    (tmp_path / "r.cpp").write_text(CPP_RENDER)
    exe = tmp_path / "r"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{host}", f"-I{os.path.join(root, 'include')}", str(tmp_path / "r.cpp"), "-o", str(exe),
                        f"-L{lib}", "-lrtb200", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path), str(tmp_path / "frame.bin"), str(tmp_path / "frame2.bin"), str(tmp_path / "render.bmp"),
                        str(tmp_path / "frame3.bin"), str(tmp_path / "frame4.bin")], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    cpp = np.fromfile(tmp_path / "frame.bin", np.float32).reshape(256, 256, 3)
    sc = rtb.load_obj(str(tmp_path / "custom.obj"))
    sc.point_lights = np.array([[-1, 1, -1, 1, 1, 1]], np.float32)
    gpu_ctx.upload_scene(sc, rtb.BVH_LBVH_DEVICE)
    rgb, ids, t, st = gpu_ctx.render(rtb.make_camera(), rtb.make_params(256, 256, 3), want_ids=True)
    assert cpp.max() > 0.1
    assert np.abs(cpp - rgb).max() <= 1e-6   # same kernels; only the order of the float atomics may differ
    hit, tri, tt, rays = r.stdout.split()
    # centre ray: NDC (0,0) is the corner of pixel (128,128), stored at row 256-1-128
    assert int(tri) == ids[127, 128] and int(hit) == int(ids[127, 128] >= 0)
    assert np.float32(tt) == t[127, 128]
    assert int(rays) == st.rays
    # second frame: post-processed on the device inside renderRayTracing; the BMP is the doubly bloomed frame, clamped and truncated
    import oracle
    port = oracle.Oracle("port")
    sc.point_lights = np.array([[-1, 1, -1, 6, 6, 6]], np.float32)
    gpu_ctx.upload_scene(sc, rtb.BVH_LBVH_DEVICE)
    bright, _, _, _ = gpu_ctx.render(rtb.make_camera(), rtb.make_params(256, 256, 3), want_ids=True)
    cfg = dict(filtering_option=2, kernel=1, filter_size=3, sigma=1.5)
    want = port.postprocess(bright, **cfg)
    cpp2 = np.fromfile(tmp_path / "frame2.bin", np.float32).reshape(256, 256, 3)
    assert np.abs(want - bright).max() > 0.05          # the bloom did something
    assert np.abs(cpp2 - want).max() <= 1e-6
    _, rgba = port.postprocess(cpp2, via_write_bitmap=True, **cfg)
    bmp = np.fromfile(tmp_path / "render.bmp", np.uint8)
    assert bmp[:2].tobytes() == b"BM" and len(bmp) == 54 + 256 * 256 * 4
    px = bmp[54:].reshape(256, 256, 4)[::-1]           # bottom-up rows, BGRA
    assert np.array_equal(px[..., [2, 1, 0, 3]], rgba)
    # third frame: the OBJ's texture (map_Kd, decoded by the host importer) through the device sampler, against the CPU port
    assert len(sc.textures) == 1 and sc.textures[0].shape == (8, 8, 3) and list(sc.mesh_tex).count(0) >= 1
    sc.point_lights = np.array([[-1, 1, -1, 1, 1, 1]], np.float32)
    gpu_ctx.upload_scene(sc, rtb.BVH_LBVH_DEVICE)
    gpu_ctx.set_texturing(rtb.TEX_BILINEAR, rtb.OOB_REPEAT, rtb.OOB_CLAMP)
    try:
        tex_py, _, _, _ = gpu_ctx.render(rtb.make_camera(), rtb.make_params(256, 256, 3), want_ids=True)
    finally:
        gpu_ctx.set_texturing(None)
    cpp3 = np.fromfile(tmp_path / "frame3.bin", np.float32).reshape(256, 256, 3)
    assert np.abs(cpp3 - tex_py).max() <= 1e-6
    assert np.abs(cpp3 - cpp).max() > 0.1               # the texture is visible
    port.set_textures(sc.uv, sc.textures, sc.mesh_tex, oracle.TEX_BILINEAR, oracle.OOB_REPEAT, oracle.OOB_CLAMP)
    try:
        o_rgb, _, _, _ = port.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, None, rtb.make_camera(), 256, 256, max_level=3, shadow_exhaustive=True)
    finally:
        port.set_textures()
    assert np.abs(cpp3 - o_rgb).max() <= COLOUR_TOL
    # fourth frame: renderRayTracing(..., textureDebugging = true): texels only, white where the material has no texture
    port.set_textures(sc.uv, sc.textures, sc.mesh_tex, oracle.TEX_BILINEAR, oracle.OOB_REPEAT, oracle.OOB_CLAMP, use_textures=False)
    try:
        d_rgb, d_ids, _, d_st = port.render(sc.pos, sc.nrm, sc.mesh_id, sc.mats, sc.point_lights, None, rtb.make_camera(), 256, 256, max_level=3, texture_debug=True)
    finally:
        port.set_textures()
    cpp4 = np.fromfile(tmp_path / "frame4.bin", np.float32).reshape(256, 256, 3)
    assert np.abs(cpp4 - d_rgb).max() <= 1e-6
    assert (cpp4[d_ids < 0] == 0).all() and (cpp4.reshape(-1, 3) == 1).all(axis=1).sum() > 100 and d_st.shadow_queries == 0


def test_host_bands_do_not_depend_on_their_order(rtb, monkeypatch):
    """rt_render cuts a large frame into bands that leave for the host one by one, rendered from the outside in by several lanes
    with smaller traversal grids.  The frame must not depend on any of that: bands top to bottom with full grids, the default,
    and one single batch give the same ids, t, ray counts and (up to the order of the float atomics) pixels."""
    g = Golden("cornell_c1_256")
    w, h = 640, 512          # 32 tile rows
    frames = {}
    for name, env, shape in (("one batch", {}, (1, 1)), ("top to bottom", {"RTB200_BAND_ORDER": "0", "RTB200_GRID_MULT": "8"}, (3, 4)), ("default", {}, (3, 4)),
                             ("seven bands", {}, (4, 7))):
        for k in ("RTB200_BAND_ORDER", "RTB200_GRID_MULT"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = rtb.Context(0)
        try:
            ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
            ctx.set_pipeline(shape[0], shape[1], 1 << 12)
            rgb, ids, t, st = ctx.render(g.camera(), rtb.make_params(w, h, 3), want_ids=True)
            frames[name] = (rgb, ids, t, (st.primary_rays, st.shadow_queries, st.secondary_rays), st.batches)
        finally:
            ctx.close()
    ref = frames["one batch"]
    assert (ref[1] >= 0).any() and ref[3][0] == w * h and ref[4] == 1
    for name in ("top to bottom", "default", "seven bands"):
        got = frames[name]
        assert got[4] == (7 if name == "seven bands" else 4), name
        assert np.array_equal(got[1], ref[1]) and bits_equal(got[2], ref[2]), name
        assert np.abs(got[0] - ref[0]).max() <= 1e-6, name
        assert got[3] == ref[3], name


def test_render_shard_fills_a_shared_host_image(rtb):
    """rt_render_shard: every rank stores the tiles it owns straight into one page-locked host image (in a multi-process job a
    shared-memory segment each rank registers; here three contexts on one GPU and one pinned buffer).  Untouched pixels keep
    what they held, the union is bit for bit the frame one context renders on its own, ray counts add up, and pageable memory
    is refused."""
    import torch
    g = Golden("cornell_c1_256")
    w, h = 200, 150      # ragged tiles, and a row pitch (2400 bytes) that allows the 16-byte stores only on every other row group
    cam, prm = g.camera(), rtb.make_params(w, h, 3)
    whole = rtb.Context(0)
    try:
        whole.upload_scene(g.scene, rtb.BVH_SAH_HOST)
        want, _, _, st_whole = whole.render(cam, prm)
    finally:
        whole.close()
    image = torch.full((h, w, 3), -7.0, dtype=torch.float32).pin_memory()
    world, counts = 3, [0, 0, 0]
    for rank in range(world):
        ctx = rtb.Context(0)
        try:
            ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
            ctx.set_shard(rank, world)
            before = image.numpy().copy()
            st = ctx.render_shard_host(cam, prm, image.data_ptr())
            after = image.numpy()
            changed = (after != before).any(axis=2)
            # only pixels of this rank's 32x16 tiles may change (Screen rows are flipped: row r holds y = h - 1 - r)
            ys, xs = np.nonzero(changed)
            assert len(ys) > 0 and (rtb.owner_map(w, h, world)[ys, xs] == rank).all()
            for k, v in enumerate((st.primary_rays, st.shadow_queries, st.secondary_rays)):
                counts[k] += v
            with pytest.raises(rtb.RtError):
                ctx.render_shard_host(cam, prm, np.zeros((h, w, 3), np.float32).ctypes.data)   # pageable memory
        finally:
            ctx.close()
    got = image.numpy()
    assert (got != -7.0).all()
    assert np.abs(got - want).max() <= 1e-6
    assert tuple(counts) == (st_whole.primary_rays, st_whole.shadow_queries, st_whole.secondary_rays)


@pytest.mark.parametrize("size,shape", [((640, 512), (2, 2)), ((200, 150), (1, 1)), ((1000, 333), (3, 5))])
def test_render_into_pinned_host_memory_equals_pageable(rtb, size, shape):
    """rt_render into page-locked memory stores the frame from the kernels themselves (background rows right after level 0, the
    rows with hits when their batch is resolved); into pageable memory it stages bands through the copy engine.  Same frame
    either way, every pixel written (the buffer starts as garbage), ids and ray counts equal."""
    import torch
    g = Golden("cornell_c1_256")
    w, h = size
    ctx = rtb.Context(0)
    try:
        ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
        ctx.set_pipeline(shape[0], shape[1], 1 << 12)
        for k, cam in enumerate((rtb.make_camera(dist=4.5), g.camera(), rtb.make_camera(euler_deg=(70.0, 200.0, 0.0), dist=6.0))):
            # one sample per pixel, and (automatic pipeline only) the 4-tap and 16-sample modes: a pixel is background when all its samples miss
            mode = (0, 1, 2)[k] if shape == (1, 1) else 0
            if mode:
                ctx.set_pipeline(0, 1)
            prm = rtb.make_params(w, h, 3, sample_mode=mode, sample_size=16)
            want, ids, t, st = ctx.render(cam, prm, want_ids=True)
            pinned = torch.full((h, w, 3), float("nan"), dtype=torch.float32).pin_memory()
            st2 = ctx.render_host_ptr(cam, prm, pinned.data_ptr())
            got = pinned.numpy()
            assert np.isfinite(got).all()
            assert np.abs(got - want).max() <= 1e-6
            assert (got[ids >= 0].sum(axis=1) > 0).any()
            if mode == 0:   # (with several samples the ids are those of the pixel's first sample only)
                assert (got[ids < 0] == 0).all()
            assert (st2.primary_rays, st2.shadow_queries, st2.secondary_rays) == (st.primary_rays, st.shadow_queries, st.secondary_rays)
    finally:
        ctx.close()
