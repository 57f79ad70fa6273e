"""CPU suite: the C-ABI library is built, loads, exports every symbol include/pst_b200.h
declares, and fails loudly (no CPU fallback) when no GPU is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pst_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pyseistr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pst_b200.h but not exported"


def test_ctypes_table_covers_header(lib):
    from pyseistr_b200 import _lib
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.pst_version()
    assert isinstance(lib.pst_last_error(), bytes)


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a GPU the product path must raise, not compute on the CPU."""
    if lib.pst_device_count() > 0:
        pytest.skip("a GPU is visible here")
    import pyseistr_b200 as ps
    with pytest.raises(ps.PstError) as e:
        ps.dip3dc(np.zeros((16, 4, 3), np.float32), verb=0)
    assert e.value.code == -2
    with pytest.raises(ps.PstError):
        ps.somf3dc(np.zeros((16, 4, 3), np.float32), np.zeros((16, 4, 3)), np.zeros((16, 4, 3)), 2, 2, 0.01, 2)


def test_product_path_does_not_import_oracle():
    """The shipped package must not reference oracle/ (the checker) anywhere."""
    pkg = os.path.join(ROOT, "pyseistr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, f
