import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle", ROOT / "feastkit.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_lib():
    """Every test session starts from current in-tree libraries (libfeastcuda.so and the example operators); build() is a no-op
    when they are newer than their sources."""
    import __graft_entry__ as g
    g.build()
    return g.LIB


@pytest.fixture(scope="session")
def engine(built_lib):
    import feastcuda as fc
    return fc.default_engine(0)
