"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs in the GPU-less build container: oracle vs compiled reference,
golden vectors, host logic, ABI symbol checks.  ``-m gpu`` needs a B200 and goes
through the C ABI (libppmx_gpu.so) only.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.orc()


@pytest.fixture(scope="session")
def ref():
    import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("compiled reference (oracle/_ref) not available")
    return r
