import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference; skips when neither oracle/_ref nor /root/reference is available."""
    from oracle import bindings
    if not bindings.ref_available() and not os.path.isdir(bindings.REF_DIR):
        pytest.skip("compiled reference (oracle/_ref) not available on this box")
    return bindings.Oracle("reference")


@pytest.fixture(scope="session")
def ctx():
    from msdfgen_b200.api import Context
    return Context(0)
