"""CPU tests of the checker itself: the restated port against the golden vectors minted from the reference's own
translation units, and (where oracle/_ref was built, i.e. in the container that has /root/reference) against that
library run live."""
import numpy as np
import pytest

from util import GOLDEN_NAMES, Golden, bits_equal

FAST = [n for n in GOLDEN_NAMES if not n.startswith("dragon")]


@pytest.mark.parametrize("name", FAST)
def test_port_reproduces_golden(name):
    g = Golden(name)
    rgb, ids, t, st = g.oracle_render("port")
    assert np.array_equal(ids, g.ids)
    assert bits_equal(t, g.t)
    assert bits_equal(rgb, g.rgb), f"max diff {np.abs(rgb - g.rgb).max()}"
    assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == g.counts


@pytest.mark.parametrize("name", ["cube_96", "cornell_sph10_aa_80x48", "cube_preset_spot_128", "zfight_96"])
def test_port_cull_free_variant(name):
    g = Golden(name)
    rgb, ids, t, st = g.oracle_render("port", shadow_exhaustive=True)
    assert bits_equal(rgb, g.rgb_x)
    assert np.array_equal(ids, g.ids_x) and bits_equal(t, g.t)
    assert (st.primary_rays, st.shadow_queries, st.secondary_rays) == g.counts_x


def test_equal_t_resolve_by_visiting_order():
    """zfight_96: coplanar triangles whose order in the reference's BVH is the reverse of the mesh order.  The as-run
    reference (useBVH=true, fixture rgb) shows the first-visited one; its useBVH=false loop the lower id; the cull-free port
    reproduces the former bit for bit, so the visiting order is restated correctly."""
    g = Golden("zfight_96")
    ties = g.ids != g.ids_x
    assert ties.sum() > 500
    assert set(np.unique(g.ids[ties])) == {0} and set(np.unique(g.ids_x[ties])) == {1}
    assert bits_equal(g.rgb, g.rgb_x)
    by_id = g.oracle_render("port", use_bvh=False)
    assert np.array_equal(by_id[1], g.ids)
    assert np.abs(by_id[0] - g.rgb).max(axis=2)[ties].min() > 0.05   # red instead of green on every tie pixel


def test_reference_bvh_culling_is_rare_in_the_fixtures():
    """Pixels where the as-run reference differs from its own cull-free evaluation (its AABB test dropped a triangle the
    triangle test accepts) are excluded from the as-run comparison of the GPU tests: there must be next to none."""
    total = bad = 0
    for name in GOLDEN_NAMES:
        g = Golden(name)
        d = np.abs(g.rgb - g.rgb_x).max(axis=2) > 1e-4
        assert d.sum() <= 3, name
        bad += int(d.sum())
        total += d.size
    assert bad / total < 2e-5


def test_golden_records_port_reference_agreement():
    """Every fixture stores whether port == verbatim reference (ids, t, rgb) when it was minted: ids and t always; rgb
    wherever the reference's colour is defined (scenes that do not trip its uninitialised-barycentric read)."""
    for name in GOLDEN_NAMES:
        g = Golden(name)
        ids_eq, t_eq, rgb_eq = (bool(v) for v in g.d["port_equals_reference"])
        assert ids_eq and t_eq, name
        if str(g.d["colour_from"]) == "reference":
            assert rgb_eq, name


def test_port_against_reference_library_live():
    import oracle
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    g = Golden("cornell_inside_128")
    a = g.oracle_render("reference")
    b = g.oracle_render("port")
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2]) and bits_equal(a[0], b[0])
    assert a[3].rays == b[3].rays


def test_brute_force_equals_bvh_in_the_oracle():
    """the reference's useBVH=false and useBVH=true paths give the same frame on the Cornell box."""
    g = Golden("cornell_c1_256")
    a = g.oracle_render("port", use_bvh=True)
    b = g.oracle_render("port", use_bvh=False)
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2]) and bits_equal(a[0], b[0])


def test_closest_hit_api_and_unnormalised_directions():
    """t is measured along normalize(d) but the hit point uses d itself (ray_tracing.cpp:65,111): scaling the direction
    changes the reference's answer, and the oracle must show that quirk."""
    import oracle
    g = Golden("monkey_192")
    rng = np.random.default_rng(1)
    o = np.tile(np.array([[0.0, 0.0, -3.0]], np.float32), (2000, 1))
    tgt = rng.uniform(-0.5, 0.5, (2000, 3)).astype(np.float32)
    d = tgt - o
    unit = d / np.linalg.norm(d, axis=1, keepdims=True)
    P = oracle.Oracle("port")
    ids1, t1 = P.closest_hit(g.scene.pos, g.scene.nrm, g.scene.mesh_id, np.concatenate([o, unit], 1))
    ids3, t3 = P.closest_hit(g.scene.pos, g.scene.nrm, g.scene.mesh_id, np.concatenate([o, 3 * unit], 1))
    assert (ids1 >= 0).sum() > 500
    assert (ids1 != ids3).sum() > 0  # the un-normalised rays land elsewhere


def test_port_reproduces_screen_postprocessing_golden():
    """tests/golden/post_screen_56x40.npz holds what the reference's own Screen (src/screen.cpp, verbatim) makes of one
    HDR image under 11 settings, through postprocessImage and through writeBitmapToFile: the port must match bit for bit."""
    import os
    import oracle
    from util import GOLDEN, POST_CONFIGS
    d = np.load(os.path.join(GOLDEN, "post_screen_56x40.npz"))
    port = oracle.Oracle("port")
    for k, cfg in enumerate(POST_CONFIGS):
        a = port.postprocess(d["img"], **cfg)
        b, rgba = port.postprocess(d["img"], via_write_bitmap=True, **cfg)
        assert bits_equal(a, d[f"post_{k}"]), cfg
        assert bits_equal(b, d[f"bmp_{k}"]) and np.array_equal(rgba, d[f"rgba_{k}"]), cfg


def test_port_postprocessing_against_reference_library_live():
    import oracle
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    from util import POST_CONFIGS
    rng = np.random.default_rng(11)
    img = (rng.random((33, 47, 3), dtype=np.float32) * 3.0).astype(np.float32)
    ref, port = oracle.Oracle("reference"), oracle.Oracle("port")
    for cfg in POST_CONFIGS:
        assert bits_equal(ref.postprocess(img, **cfg), port.postprocess(img, **cfg)), cfg
        (a, ra), (b, rb) = ref.postprocess(img, via_write_bitmap=True, **cfg), port.postprocess(img, via_write_bitmap=True, **cfg)
        assert bits_equal(a, b) and np.array_equal(ra, rb), cfg
