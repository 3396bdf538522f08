import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle_f32():
    from oracle import oracle
    return oracle.Oracle("f32")


@pytest.fixture(scope="session")
def oracle_f64():
    from oracle import oracle
    return oracle.Oracle("f64")


@pytest.fixture(scope="session", params=["f32", "f64"])
def oracle_any(request):
    from oracle import oracle
    return oracle.Oracle(request.param)
