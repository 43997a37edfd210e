import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))



def _ensure_built():
    """The CPU suite needs libppm_b200.so (host functions + ABI surface) and the oracle: build them if absent."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "ppmpa_b200", "libppm_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ppmpa_b200", "csrc"), "-s", "all"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "libppm_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


_ensure_built()


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
