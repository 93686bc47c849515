"""CPU: the C-ABI library loads and exports every symbol include/traffic_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(REPO, "include", "traffic_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tb2_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("tb2_create", "tb2_destroy", "tb2_upload", "tb2_download", "tb2_prepare", "tb2_step", "tb2_step_host"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from traffic_b200 import native
    native.build()
    L = ctypes.CDLL(native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), "libtraffic_b200.so does not export " + s
    assert sorted(native.EXPORTS) == declared_symbols(), "binding list and header disagree"
    L.tb2_abi_version.restype = ctypes.c_int
    assert L.tb2_abi_version() == native.ABI_VERSION


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from traffic_b200 import native
    from traffic_b200.state import make_world
    w = make_world((0.0, 1000.0, 0.0, 1000.0), np.zeros((0, 2)), np.zeros(1, np.int64), np.zeros((0, 2)), np.zeros((0, 2)))
    with pytest.raises(native.NativeError):
        native.DeviceSim(w)
    # and through the raw C ABI
    cfg = native.config_of(w)
    h = ctypes.c_void_p()
    rc = native.lib().tb2_create(ctypes.byref(cfg), None, 0, ctypes.byref(h))
    assert rc == 3, "expected TB2_ERR_NO_DEVICE"


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "traffic_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "traffic_oracle" not in txt, f
