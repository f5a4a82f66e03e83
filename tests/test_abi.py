"""The C-ABI boundary: libamh_b200.so loads without a GPU, exports every symbol include/amh.h declares,
and fails LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "advancedmh.jl_b200", "libamh_b200.so")


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "amh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amh_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points(amh):
    names = _declared_symbols()
    assert len(names) >= 21
    from advancedmh_jl_b200 import _capi
    assert sorted("amh_" + n for n in _capi.ABI_SYMBOLS) == names


def test_product_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(LIB)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/amh.h but not exported"
    major, minor = C.c_int32(), C.c_int32()
    assert lib.amh_version(C.byref(major), C.byref(minor)) == 0
    assert (major.value, minor.value) == (0, 1)
    assert lib.amh_contract_version() == 2


def test_oracle_exports_the_same_abi_under_its_own_prefix(oracle):
    for name in _declared_symbols():
        assert hasattr(oracle.lib, "amho_" + name[4:])


def test_no_cpu_fallback_without_a_device(amh):
    """on a box without a GPU the product path must raise, never compute"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(amh.AMHError, match="no CUDA device"):
        amh.Engine()
    model = amh.DensityModel(amh.MvNormalTarget(None, [[1.0]]))
    with pytest.raises(amh.AMHError, match="no CUDA device"):
        amh.sample(model, amh.RWMH(1), 10)


def test_product_package_never_references_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (prompt rule 3)"""
    pkg = os.path.join(ROOT, "advancedmh.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libamh_oracle" not in text and "amho_" not in text.replace("``amho_``", ""), f
