"""GPU parity tests of the Screen post-processing: rt_postprocess / rt_set_postprocess through the C ABI
against the golden vectors minted from the reference's own Screen and against the CPU port run live.

Bars: bloom (box and Gaussian), clamp, Reinhard map and the 8-bit conversion are bit-exact; the exposure map and the gamma
curve involve exp / pow (glibc on the CPU, CUDA's double-precision functions rounded to float on the device): 2e-6 relative.
"""
import os

import numpy as np
import pytest

from util import GOLDEN, POST_CONFIGS, POST_EXACT, Golden, bits_equal

pytestmark = pytest.mark.gpu

TRANSCENDENTAL_RTOL = 2e-6


def _post(rtb, cfg):
    return rtb.make_post(**cfg)


def _close(a, b):
    return np.allclose(a, b, rtol=TRANSCENDENTAL_RTOL, atol=1e-7, equal_nan=True)


def test_postprocess_matches_reference_screen_golden(rtb, gpu_ctx):
    d = np.load(os.path.join(GOLDEN, "post_screen_56x40.npz"))
    for k, cfg in enumerate(POST_CONFIGS):
        a = gpu_ctx.postprocess(d["img"], _post(rtb, cfg))
        b, rgba = gpu_ctx.postprocess(d["img"], _post(rtb, cfg), via_write_bitmap=True)
        if k in POST_EXACT:
            assert bits_equal(a, d[f"post_{k}"]), f"postprocessImage differs for {cfg}: max {np.abs(a - d[f'post_{k}']).max()}"
        else:
            assert _close(a, d[f"post_{k}"]), f"postprocessImage differs for {cfg}: max {np.abs(a - d[f'post_{k}']).max()}"
        if cfg["filtering_option"] != 3:   # writeBitmapToFile never applies gamma: exact unless the exposure map is involved
            assert bits_equal(b, d[f"bmp_{k}"]) and np.array_equal(rgba, d[f"rgba_{k}"]), f"writeBitmapToFile differs for {cfg}"
        else:
            assert _close(b, d[f"bmp_{k}"])
            assert np.abs(rgba.astype(int) - d[f"rgba_{k}"].astype(int)).max() <= 1   # truncation next to an integer boundary


@pytest.mark.parametrize("shape", [(45, 70), (1, 1), (16, 333), (130, 97)])
def test_postprocess_against_live_port(rtb, gpu_ctx, shape):
    """Ragged sizes (not multiples of the 32x8 block), one-pixel image, strong HDR: box / Gaussian bloom of every option."""
    import oracle
    port = oracle.Oracle("port")
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = (rng.random(shape + (3,), dtype=np.float32) ** 2 * 4.0).astype(np.float32)
    for opt in range(6):
        for kern in (0, 1):
            cfg = dict(filtering_option=opt, kernel=kern, filter_size=4, kernel_repetitions=2, sigma=1.5)
            got, want = gpu_ctx.postprocess(img, _post(rtb, cfg)), port.postprocess(img, **cfg)
            if opt == 3:
                assert _close(got, want), cfg
            else:
                assert bits_equal(got, want), f"{cfg}: max {np.abs(got - want).max()}"


def test_postprocess_inside_the_frame(rtb, gpu_ctx):
    """renderRayTracing ends with screen.postprocessImage() (main.cpp:397-398): with rt_set_postprocess the frame comes back
    post-processed, and equals the CPU reference's post-processing of the same frame rendered without it."""
    import oracle
    g = Golden("spheres_preset_160")          # point light of colour 15: plenty of pixels above brightness 1
    gpu_ctx.upload_scene(g.scene, rtb.BVH_SAH_HOST)
    gpu_ctx.set_postprocess(None)
    plain, _, _, _ = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
    assert (plain @ np.array([0.2126, 0.7152, 0.0722], np.float32) >= 1).sum() > 200
    port = oracle.Oracle("port")
    try:
        for cfg in (dict(filtering_option=2), dict(filtering_option=1, kernel=1, filter_size=3), dict(filtering_option=2, gamma_correction=True)):
            gpu_ctx.set_postprocess(_post(rtb, cfg))
            got, ids, t, st = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
            want = port.postprocess(plain, **cfg)
            if cfg.get("gamma_correction"):
                assert _close(got, want), cfg
            else:
                assert bits_equal(got, want), f"{cfg}: max {np.abs(got - want).max()}"
            assert np.array_equal(ids, g.ids_x)   # the traced frame itself is untouched
        # bloom not live and no gamma: postprocessImage does nothing
        gpu_ctx.set_postprocess(_post(rtb, dict(filtering_option=2, bloom_live=False)))
        same, _, _, _ = gpu_ctx.render(g.camera(), g.params(), want_ids=True)
        assert bits_equal(same, plain)
    finally:
        gpu_ctx.set_postprocess(None)


def test_postprocess_full_size_properties(rtb, gpu_ctx):
    """4K frame (C3's size).  The filter is local, so any window of the result equals the CPU port's result on that window
    plus a margin of filter_size * repetitions pixels; windows at two corners check the black border, one lies inside.
    Filter sizes 16 and 17 straddle the switch between the staged (shared-memory) and the direct kernel."""
    import oracle
    port = oracle.Oracle("port")
    h, w = 2160, 3840
    rng = np.random.default_rng(7)
    img = (rng.random((h, w, 3), dtype=np.float32) ** 4 * 6.0).astype(np.float32)
    for cfg, margin in ((dict(filtering_option=2, filter_size=5, kernel_repetitions=2), 10), (dict(filtering_option=1, kernel=1, filter_size=17), 17),
                        (dict(filtering_option=1, kernel=0, filter_size=16), 16)):
        got = gpu_ctx.postprocess(img, _post(rtb, cfg))
        for (y0, x0) in ((0, 0), (h - 96, w - 128), (1000, 2000)):          # two corners (border handling) and the interior
            ys, xs = slice(max(0, y0 - margin), min(h, y0 + 96 + margin)), slice(max(0, x0 - margin), min(w, x0 + 128 + margin))
            want = port.postprocess(img[ys, xs], **cfg)
            oy, ox = y0 - ys.start, x0 - xs.start
            assert bits_equal(got[y0:y0 + 96, x0:x0 + 128], want[oy:oy + 96, ox:ox + 128]), (cfg, y0, x0)


def test_postprocess_errors_are_loud(rtb, gpu_ctx):
    img = np.zeros((4, 4, 3), np.float32)
    with pytest.raises(rtb.RtError):
        gpu_ctx.postprocess(img, rtb.make_post(filtering_option=9))
    with pytest.raises(rtb.RtError):
        gpu_ctx.postprocess(img, rtb.make_post(filtering_option=1, filter_size=100))
