import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REF_EXAMPLES = os.path.join(ROOT, "tests", "golden", "scenes")   # committed copies of the *data* files used by tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def engine():
    """One engine for the whole GPU session.  No fallback: fails loudly without a GPU."""
    import ppmpa_b200
    eng = ppmpa_b200.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def builtin_scene():
    import ppmpa_b200
    return ppmpa_b200.read_scene()
