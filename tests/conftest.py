import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One engine context for the whole GPU session; fails loudly when the engine cannot start."""
    from keras_unsupervised_b200.engine import Context

    return Context(device=0, seed=42)
