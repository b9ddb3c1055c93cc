import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (if needed) the oracle and the product library once per session."""
    import shutil
    import __graft_entry__ as g
    so = os.path.join(ROOT, "subrosadg_b200", "libsubrosadg_b200.so")
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc") and not os.path.exists(so):
        pytest.skip("neither the CUDA toolkit (nvcc) nor a prebuilt libsubrosadg_b200.so is available on this machine")
    g.build()
    return True
