"""TEST INFRASTRUCTURE — mints tests/golden/post_screen_56x40.npz from the reference's own Screen (src/screen.cpp compiled
verbatim into oracle/_ref, see oracle/Makefile).  Run in the container that has /root/reference:
    python tests/golden/make_golden_post.py
Stores one HDR input image and, for a list of Screen settings, what Screen::postprocessImage leaves in the pixels and what
Screen::writeBitmapToFile hands to its BMP encoder.  The restated port must reproduce every entry bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from util import POST_CONFIGS  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    img = (rng.random((40, 56, 3), dtype=np.float32) ** 3 * 2.5).astype(np.float32)
    img[5:9, 10:14] = [4.0, 3.0, 0.5]       # a saturated patch
    img[30:33, 50:56] = [0.0, 6.0, 0.0]     # one touching the right border
    img[0, 0] = [9.0, 9.0, 9.0]             # and a corner
    ref, port = oracle.Oracle("reference"), oracle.Oracle("port")
    out = dict(img=img)
    for k, cfg in enumerate(POST_CONFIGS):
        a = ref.postprocess(img, **cfg)
        b, rgba = ref.postprocess(img, via_write_bitmap=True, **cfg)
        pa = port.postprocess(img, **cfg)
        pb, prgba = port.postprocess(img, via_write_bitmap=True, **cfg)
        assert np.array_equal(a.view(np.int32), pa.view(np.int32)), cfg
        assert np.array_equal(b.view(np.int32), pb.view(np.int32)) and np.array_equal(rgba, prgba), cfg
        out[f"post_{k}"], out[f"bmp_{k}"], out[f"rgba_{k}"] = a, b, rgba
        print(k, cfg, "changed pixels:", int((np.abs(a - img).max(axis=2) > 0).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "post_screen_56x40.npz"), **out)


if __name__ == "__main__":
    main()
