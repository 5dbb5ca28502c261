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
    """Make sure the host-side native pieces exist (scene generator, oracle).  Never builds CUDA here."""
    from stratusgfx_b200.scene import build as build_scene
    from oracle.oracle_py import build as build_oracle
    build_scene()
    build_oracle()
    return True


@pytest.fixture(scope="session")
def tiny(built):
    from tests.parity_util import SceneData
    return SceneData("tiny", frames=2)


@pytest.fixture(scope="session")
def tiny_oracle(tiny):
    from tests.parity_util import run_oracle
    return run_oracle(tiny)
