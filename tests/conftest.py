import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    from oracle import dbg_oracle
    dbg_oracle.build()
    return dbg_oracle


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session.  No fallback: if the library or the GPU is missing this raises."""
    import rogtk_b200
    eng = rogtk_b200.Engine(device=0)
    yield eng
    eng.close()
