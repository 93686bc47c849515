import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _bounds_check_build(request):
    """With a -DTB2_DEBUG_BOUNDS library (scripts/gpu_bounds.sh: TB2_LIB=build_variants/lib_bounds.so) every GPU test also
    asserts that no step kernel was about to dereference an out-of-range index (compute-sanitizer is closed on the pool)."""
    yield
    if "gpu" not in request.keywords or not os.environ.get("TB2_LIB"):
        return
    try:
        from traffic_b200 import native
        L = native.lib()
    except Exception:
        return
    if L.tb2_debug_bounds_enabled():
        v = L.tb2_debug_bounds_violations()
        assert v == 0, "bounds rules violated (mask 0x%x) in %s" % (v, request.node.name)
