import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-group27_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# Tests marked `gpu` are never skipped on a box without a GPU: rtb200.Context(0) raises there (the CUDA path is the product, there
# is no CPU fallback), so `-m gpu` on such a box fails loudly instead of passing on nothing.


@pytest.fixture(scope="session")
def rtb():
    import rtb200
    return rtb200


@pytest.fixture(scope="session")
def gpu_ctx(rtb):
    ctx = rtb.Context(0)  # raises without a GPU: no fallback
    yield ctx
    ctx.close()
