"""CPU: the C-ABI library loads and exports every symbol include/synthpy_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "synthpy_b200.h")
LIB = os.path.join(ROOT, "synthpy_b200", "csrc", "libsynthpy_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header():
    from synthpy_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib.sp_version() == 1


def test_struct_sizes_match_header_layout():
    from synthpy_b200 import _lib as L
    assert ctypes.sizeof(L.OpticOp) == 32 and ctypes.sizeof(L.Beam) == 48
    assert ctypes.sizeof(L.Image) == 64 and ctypes.sizeof(L.Channel) == 88
    assert ctypes.sizeof(L.Params) == 80 and ctypes.sizeof(L.Stats) == 48


def test_no_cpu_fallback():
    """Product code never imports the oracle, and compute entry points refuse to run without CUDA."""
    import torch
    pkg = os.path.join(ROOT, "synthpy_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M) and "oracle" not in src, fn
    if not torch.cuda.is_available():
        from synthpy_b200 import domain, propagator
        import numpy as np
        dom = domain.ScalarDomain([1e-2, 1e-2, 2e-2], 8, ne_type="test_null")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            propagator.solve(np.zeros((9, 4)), dom, 1e-2)
