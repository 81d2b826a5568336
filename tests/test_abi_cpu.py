"""CPU suite: the C-ABI library loads without a GPU and exports every symbol include/repsurf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "repsurf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from repsurf_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 22
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/repsurf_b200.h but not exported"
    lib.rsb_abi_version.restype = ctypes.c_int
    assert lib.rsb_abi_version() >= 1
    # and the python binding table covers the same set
    assert set(_native.EXPORTS) == set(names)


def test_product_has_no_cpu_fallback():
    import torch
    from repsurf_b200.cls import pointops as P
    with pytest.raises(RuntimeError):
        P.furthestsampling(torch.rand(1, 16, 3), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "repsurf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
