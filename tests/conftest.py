import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracles
    return oracles.load_oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import oracles
    lib = oracles.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libpg_ref.so not built (needs the reference tree; `make ref` in the build container)")
    return lib


@pytest.fixture(scope="session")
def engine():
    import pangenie_b200 as pg
    return pg.Engine(0)
