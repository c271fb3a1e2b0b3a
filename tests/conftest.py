import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# keep the Monte-Carlo set-up loops (voronoi Lloyd relaxation, subvolume volumes) short in the test-suite
os.environ.setdefault("NK_VORONOI_MAX_SAMPLES", "20000")
os.environ.setdefault("NK_VOLUME_MAX_SAMPLES", "50000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _library_built():
    """The C-ABI library is a build artefact (git-ignored): compile it on first use in a fresh checkout
    (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "nanokappa_b200", "libnk_b200.so")
    if not os.path.isfile(lib):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
