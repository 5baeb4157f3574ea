import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and the product libraries once per session (no-ops when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def gpu_ctx():
    import woxel_b200 as W
    import knobs
    ctx = knobs.apply_env(W.Context())  # raises WxError(WX_ERR_NO_DEVICE) without a GPU: there is no CPU fallback
    yield ctx
    ctx.close()
