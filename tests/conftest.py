import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import oracle as O
    O.build(("port",))
    return O.port()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import oracle as O
    if os.path.isdir("/root/reference/src"):
        O.build(("ref",))
    if not O.have_ref():
        pytest.skip("oracle/_ref/libbbo_ref.so absent (reference tree not available)")
    return O.ref()
