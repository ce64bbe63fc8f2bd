import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    return oracle.Oracle()


@pytest.fixture(scope="session")
def reference_lib():
    import oracle

    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libsatsuma_ref.so not built (needs /root/reference)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def golden_samples():
    return np.load(os.path.join(GOLDEN, "samples.npz"))


@pytest.fixture(scope="session")
def golden_synthetic():
    return np.load(os.path.join(GOLDEN, "synthetic.npz"))


@pytest.fixture(scope="session")
def sx():
    """The product package with its CUDA library built (build() cross-compiles without a GPU)."""
    from satsuma2_b200 import build as sxbuild

    sxbuild.build()
    import satsuma2_b200

    satsuma2_b200.load_library()
    return satsuma2_b200
