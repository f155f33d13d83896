import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def scene():
    from scene_util import small_scene
    return small_scene()


@pytest.fixture(scope="session")
def oracle(scene):
    from oracle.bindings import OracleLib
    return OracleLib.from_scene(scene)


@pytest.fixture(scope="session")
def gpu(pkg, scene):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = pkg.PmvsB200.from_scene(scene)
    yield g
    g.close()
