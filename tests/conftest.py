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
    # The checked build (RTB200_LIB=.../librtb200_checked.so, include/rt_b200.h: rt_violations) counts every index its kernels found
    # out of range, in all contexts of this process: a session run against it ends with the verdict.  The default build reports zeros.
    verdict = {"checked": rtb.checked_build(), "violations": ctx.violations()}
    ctx.violations_selftest()  # the counters are alive: two deliberate violations show up (checked build), nothing changes (default build)
    after = ctx.violations()
    verdict["selftest"] = {k: after[k] - verdict["violations"][k] for k in after if after[k] != verdict["violations"][k]}
    ctx.close()
    out = os.environ.get("RTB200_VIOLATIONS_OUT")
    if out:
        import json
        with open(out, "w") as f:
            json.dump(verdict, f)
    assert verdict["selftest"] == ({"table_entry": 1, "wide_stack_slot": 1} if verdict["checked"] else {}), verdict
    assert not any(verdict["violations"].values()), f"out-of-range indices in the kernels: {verdict}"
