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
