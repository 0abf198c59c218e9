"""pytest configuration: the `gpu` marker and the shared backend fixture."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA backend (nvcc cross-compiles without a GPU) and the CPU oracle."""
    from segalign_b200 import build
    build.build_backend()
    build.build_oracle()
    return build


@pytest.fixture()
def backend(built):
    """A freshly initialised backend on GPU 0.  No fallback: fails loudly without a device."""
    from segalign_b200.backend import Backend
    be = Backend()
    be.InitializeInterface(1)
    yield be
    try:
        be.ShutdownProcessor()
    except Exception:
        pass
