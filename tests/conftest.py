import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def vb():
    """The product binding (ctypes over libviterbi_b200.so)."""
    import viterbi_dll_b200

    return viterbi_dll_b200


@pytest.fixture(scope="session")
def port():
    import oracle_lib

    return oracle_lib.port()


@pytest.fixture(scope="session")
def checker():
    """Compiled reference when present (oracle/_ref), else the C port."""
    import oracle_lib

    return oracle_lib.checker()
