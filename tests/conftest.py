import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pa():
    return importlib.import_module("privacy-auction_b200")


@pytest.fixture(scope="session")
def engine(pa):
    """The CUDA engine on cuda:0.  No fallback: if the library or the GPU is
    missing the gpu tests fail."""
    eng = pa.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()
