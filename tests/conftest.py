import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lmpc():
    import lmpc_b200
    if not os.path.exists(lmpc_b200.lib_path()):
        # fresh checkout (the .so is git-ignored): build it the way the driver does
        import __graft_entry__
        __graft_entry__.build()
    lmpc_b200.load_library()
    return lmpc_b200
