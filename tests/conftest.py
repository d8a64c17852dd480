import os
import sys

import pytest

# tests/test_gpu_comm.py runs several ranks of the peer-memory protocol as threads of THIS process on one device: while one
# rank's stream spins in an on-device wait, another rank's first launch of a kernel must not need the device to go idle.
# Lazy module loading does (loading code is an allocation, an implicit synchronisation point), so load everything up front.
# (One process per GPU — the way the library is deployed and bench.py runs — is not affected.)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

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
