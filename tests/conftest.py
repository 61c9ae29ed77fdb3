import json
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("optimal-py-learning_b200", "oracle", "tests"):
    path = os.path.join(ROOT, sub)
    if path not in sys.path:
        sys.path.insert(0, path)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "meta.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_objgrad():
    return numpy.load(os.path.join(GOLDEN, "objgrad.npz"))


@pytest.fixture(scope="session")
def golden_qn():
    return numpy.load(os.path.join(GOLDEN, "quasi_newton.npz"))


@pytest.fixture(scope="session")
def golden_traces():
    return numpy.load(os.path.join(GOLDEN, "traces.npz"))
