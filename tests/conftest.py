import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure both shared libraries exist (the product .so and the oracle .so)."""
    import oracle
    from rodio_b200 import build as rb_build
    oracle.build()
    rb_build.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    import rodio_b200 as rb
    return rb.default_context(0)
