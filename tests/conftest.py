import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def B():
    """The product package (ctypes binding over libbaorec_b200.so)."""
    import __graft_entry__ as G
    lib = G.PKG_DIR / "lib" / "libbaorec_b200.so"
    if not lib.exists():
        G.build()
    return G.load_package()


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import baorec_oracle
    return baorec_oracle
