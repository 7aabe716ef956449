"""CPU test: libwfst_b200.so loads and exports every symbol include/wfst_b200.h
declares (no compute calls — there is no GPU here), and the product never imports
the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wfst_b200.h")).read()
    return sorted(set(re.findall(r"WFST_API[^;(]*?\b(wfst_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from gtn_applications_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 10
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes binding for " + n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().wfst_abi_version() == 2
    assert _lib.lib().wfst_last_error() == b""


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gtn_applications_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(gtn|gtn64|ref_criterions|dp_numpy|_gtn_oracle)\b",
                                     text, re.M), f
                assert "oracle/" not in text.replace("the oracle", ""), f
