"""CPU tier: the CUDA library loads and exports every symbol include/ttipm.h declares (no compute calls),
and the ctypes binding declares the same set."""
import ctypes
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "tensor-train-interior-point-method_b200")


def header_functions():
    src = open(os.path.join(ROOT, "include", "ttipm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttipm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ttipm_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(mod.build())
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ttipm.h but not exported"


def test_binding_matches_header():
    import sys
    sys.path.insert(0, PKG)
    from ttipm_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == header_functions()


def test_product_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import sys
    sys.path.insert(0, PKG)
    from ttipm_b200.runtime import Runtime, TTIPMError
    with pytest.raises(TTIPMError):
        Runtime()
