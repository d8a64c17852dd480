import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib

    oracle_lib.build()
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def ctx():
    """One ivx_ctx for the whole GPU session. Fails loudly if the CUDA library is missing."""
    from impact_b200.voxel import Context

    c = Context(0)
    yield c
    c.close()
