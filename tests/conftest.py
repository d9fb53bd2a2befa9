import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build

    build(ref=False)
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled unmodified reference; tests that use it skip when it is absent."""
    from oracle.oracle import Ref

    if not Ref.available():
        pytest.skip("oracle/_ref/libads_ref.so not available")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(d, name + ".npz")) for name in ("setup", "solve", "problems", "flow")}
