import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
if str(HERE) not in sys.path:
    sys.path.insert(0, str(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.Oracle()


@pytest.fixture(scope="session")
def gen():
    import helpers
    return helpers.Generator()


@pytest.fixture(scope="session")
def bwts():
    """The product: ctypes mirror over libbwts_b200.so (fails loudly if absent)."""
    import helpers
    return helpers.load_product()
