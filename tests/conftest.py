import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def nbs():
    return importlib.import_module("openmm-nonbonded-slicing_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as module
    if not module.available("port"):
        module.build(("port",))
    return module
