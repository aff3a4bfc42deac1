import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return {k: dict(np.load(os.path.join(g, k + ".npz"))) for k in ("board", "mcts", "net", "net20", "selfplay")
            if os.path.exists(os.path.join(g, k + ".npz"))}
