import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `pytest -m gpu`")


@pytest.fixture(scope="session", autouse=True)
def _build_library():
    """The C-ABI library must exist for both suites (the CPU suite checks its exports)."""
    import __graft_entry__ as g
    g.build_if_stale()
