"""pytest configuration: registers the `gpu` marker and puts `oracle/` (test
infrastructure) on sys.path for the tests only."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
REFERENCE = "/root/reference"

for p in (ROOT, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line(
        "markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(os.path.join(REFERENCE, "criterions"))
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def gtn32():
    import gtn
    return gtn


@pytest.fixture(scope="session")
def gtn64():
    import gtn64
    return gtn64


@pytest.fixture(params=["lean-pair", "lean-pair-narrow", "lean-single", "generic"])
def lattice_kernel(request):
    """Runs a GPU test on each lattice kernel through the C-ABI test hook: the shared-memory
    "lean" kernel as a cluster of two blocks that meet in the middle (whenever T allows), the
    same without its wide-register variant for CSR acceptors of 1025..2048 nodes, the single-block
    lean kernel, and the generic global-memory kernel."""
    from gtn_applications_b200 import _lib
    old = _lib.lib().wfst_debug_force_generic_lattice({"lean-pair": 3, "lean-pair-narrow": 4, "lean-single": 2, "generic": 1}[request.param])
    yield request.param
    _lib.lib().wfst_debug_force_generic_lattice(old)
