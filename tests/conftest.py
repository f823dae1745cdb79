import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds the CUDA library (nvcc cross-compiles on CPU) and the oracle once per session."""
    from cam_nor_physics_b200 import build
    build.build()
    from oracle_lib import build_oracle
    build_oracle()
    return True


@pytest.fixture(scope="session")
def oracle_pm(built):
    from oracle_lib import Oracle
    from cam_nor_physics_b200 import soundings as S
    o = Oracle("pm")
    assert o.convi(o.default_params(16, 32, S.limcnv_for(32))) == 0
    return o


@pytest.fixture(scope="session")
def oracle_libm(built):
    from oracle_lib import Oracle
    from cam_nor_physics_b200 import soundings as S
    o = Oracle("libm")
    assert o.convi(o.default_params(16, 32, S.limcnv_for(32))) == 0
    return o
