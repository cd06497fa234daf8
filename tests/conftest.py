import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything under test is native: build it once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build(verbose=False)


@pytest.fixture(scope="session")
def assets():
    from functracer_b200 import scenes
    return scenes.asset_dir()
