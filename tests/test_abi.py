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
    assert _lib.lib.sp_version() == 4


def test_struct_sizes_match_header_layout():
    from synthpy_b200 import _lib as L
    assert ctypes.sizeof(L.OpticOp) == 32 and ctypes.sizeof(L.Beam) == 48
    assert ctypes.sizeof(L.Image) == 64 and ctypes.sizeof(L.Channel) == 88
    assert ctypes.sizeof(L.Params) == 88 and ctypes.sizeof(L.Stats) == 48


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
        from synthpy_b200 import config
        with pytest.raises(RuntimeError, match="CUDA device|no CPU path"):
            config.jax_init()


def test_error_codes_without_touching_the_gpu():
    """Argument validation happens before any CUDA call: every entry point returns a negative code and
    sp_last_error() explains it (no exceptions across the boundary)."""
    import ctypes as C
    from synthpy_b200 import _lib as L
    lib = L.lib
    P = L.Params(method=0, n_steps=10, h=1e-12, probing_axis=2, out_axis_a=0, out_axis_b=1)
    assert lib.sp_propagate(None, C.byref(P), None, None, None, 10, 0, None, None, None, None, None, 0, None, None) == -1
    assert b"null" in lib.sp_last_error()
    h = C.c_void_p()
    ax = (C.c_float * 2)(0.0, 1.0)
    bad = (C.c_float * 1)(0.0)
    assert lib.sp_field_create(C.byref(h), C.c_void_p(8), 1, ax, ax, bad, 2, 2, 1, 1.0, 2, 0, None) == -1
    assert b"at least 2" in lib.sp_last_error()
    assert lib.sp_field_create(C.byref(h), None, 1, ax, ax, ax, 2, 2, 2, 1.0, 2, 0, None) == -1
    assert lib.sp_field_create(C.byref(h), C.c_void_p(8), 1, ax, ax, ax, 2, 2, 2, 1.0, 7, 0, None) == -1
    assert b"march_axis" in lib.sp_last_error()
    assert lib.sp_rhs(None, None, None, 0, None, None) == -1
    assert lib.sp_optics_image(None, None, 0, None, None, None, None) == -1
    assert lib.sp_beam_generate(None, 0, 0, None, None) == -1
    assert lib.sp_image_finalize(None, None, None) == -1
    assert lib.sp_field_destroy(None) == 0 and lib.sp_workspace_destroy(None) == 0
    d = C.c_void_p(8)
    assert lib.sp_scatter_to_grid(None, None, None, 1, 0, None, 0, None, None, 1, 1, 0.0, None, None, None) == -1
    assert lib.sp_scatter_to_grid(d, d, d, 0, 3, d, 1, d, d, 4, 4, 0.0, d, d, None) == -1 and b"n_val" in lib.sp_last_error()
    assert lib.sp_scatter_to_grid(d, d, d, 1, 3, None, 1, d, d, 4, 4, 0.0, d, d, None) == -1 and b"triangulation" in lib.sp_last_error()
    assert lib.sp_fresnel_prepare(d, None, 1, 4, 4, 2, 0.4, d, None) == -1 and b"null" in lib.sp_last_error()
    assert lib.sp_fresnel_prepare(d, None, 2, 4, 4, 2, 0.4, d, None) == -1 and b"mode" in lib.sp_last_error()
    assert lib.sp_fresnel_transfer(d, 4, 4, 0.0, 1.0, 1e-6, 0.1, 0.0, None) == -1 and b"spacings" in lib.sp_last_error()
    assert lib.sp_fresnel_finish(None, 4, 4, 2, 1.0, 0.0, d, None) == -1
    import pytest
    with pytest.raises(L.SynthpyB200Error, match="synthpy_b200 error -1"):
        L.check(lib.sp_rhs(None, None, None, 0, None, None))


def test_integration_stub_matches_binding():
    """The ctypes stub printed in INTEGRATION.md declares sp_params exactly like the shipped binding."""
    from synthpy_b200 import _lib as L
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = md[md.index("class Params(C.Structure)"):md.index("def check(rc)")]
    names = re.findall(r'\("(\w+)", C\.c_(\w+)\)', block)
    assert [n for n, _ in names] == [n for n, _ in L.Params._fields_]
    stub = type("Stub", (ctypes.Structure,), {"_fields_": [(n, getattr(ctypes, "c_" + t)) for n, t in names]})
    assert ctypes.sizeof(stub) == ctypes.sizeof(L.Params)
    assert all(getattr(stub, n).offset == getattr(L.Params, n).offset for n, _ in names)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps each exported function to the reference interface it replaces."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols() if n not in md]
    assert not missing, missing
