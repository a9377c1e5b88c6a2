"""CPU-side checks of the C-ABI library: it builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpfanova_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpf_[a-z_0-9A-Z]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from gpfanova_b200 import _capi
    lib = _capi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _capi.SIGNATURES, "ctypes signature missing for %s" % s
    assert lib.gpf_version() >= 100
    assert b"sm_100a" in lib.gpf_build_info()


def test_no_cpu_fallback():
    """without a CUDA device the compute entry points must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from gpfanova_b200 import engine
    with pytest.raises(RuntimeError):
        engine.rbf_K(np.zeros((4, 1)), 0.0, 0.0)
    with pytest.raises(RuntimeError):
        engine.ChainEngine(np.zeros((4, 1)), np.zeros((4, 2)), np.ones((2, 1)), [[0]])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gpfanova_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                txt = open(os.path.join(dp, fn)).read()
                assert "gpfanova_oracle" not in txt and "oracle/" not in txt, fn
