import importlib
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pbf():
    """The product's host-side binding (loads pbf-cuda_b200/libpbf_b200.so; building it if needed)."""
    lib = os.path.join(ROOT, "pbf-cuda_b200", "libpbf_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("pbf-cuda_b200")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    _oracle.lib()
    return _oracle
