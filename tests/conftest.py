"""pytest configuration: the `gpu` marker, and import plumbing for the dotted package directory."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    p = g.load_package()
    p.build()
    return p


@pytest.fixture(scope="session")
def handle(pkg):
    """A library handle on cuda:0. Fails loudly (no skip) when the device or the .so is missing."""
    return pkg.default_handle(0)
