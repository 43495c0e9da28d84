import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """oracle/ is test infrastructure: build the C restatement (and, where
    /root/reference exists, the unmodified reference binary) before any test."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


@pytest.fixture(scope="session", autouse=True)
def _library_override():
    """MB_TEST_LIB=<path>: run the tests against another build of the C-ABI library (e.g. the `-m gpu` files against
    tests/cusim/_build/libmeryl_b200_sim.so in a container without a GPU, as far as they go through the binding)."""
    path = os.environ.get("MB_TEST_LIB")
    if path:
        import meryl_b200 as mb
        mb.LIB_PATH, mb._lib = path, None
        os.environ["LD_PRELOAD"] = path          # the command line tools of the tests bind the same symbols
    yield


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a spinning kernel) must end the process, not sit on the box until the outer limit."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))
