import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; built here with nvcc if missing (cross-compiles without a GPU)."""
    from housescan_b200 import build as hb_build

    return hb_build.build()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx(built_lib):
    """GPU context.  No fallback: if the extension or the device is missing the gpu tests fail loudly."""
    import housescan_b200 as hb

    c = hb.Context(0)
    yield c
    c.close()
