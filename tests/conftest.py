import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pr-disagg-radar-gan_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def lib():
    from rdg_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def ctx16():
    from rdg_b200.engine import Context
    c = Context(16, 1, max_chunk=1024)
    yield c
    c.close()
