import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["toeplitz_512_leaf64", "utoeplitz_300_leaf32", "gauss2d_1024_leaf64"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu)")
    sys.setrecursionlimit(20000)


@pytest.fixture(scope="session")
def built():
    """The engine must be built (CPU: nvcc cross-compiles) before any test."""
    import __graft_entry__ as g
    g.build()
    import strumpack_b200 as sb
    return sb


def have_ref():
    from oracle import ref
    return ref.available()
