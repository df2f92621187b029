import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ref_mod():
    """The compiled reference rasterizer (oracle/_ref), or skip when it was not shipped."""
    import build_ref
    try:
        return build_ref.load()
    except FileNotFoundError as e:
        pytest.skip(str(e))


def golden(name):
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing: run tests/golden/make_golden.py on a GPU box")
    return np.load(path)
