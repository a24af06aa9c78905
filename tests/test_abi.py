"""The C-ABI library loads, exports every symbol include/se_b200.h declares, and fails loudly
without a device (no CPU fallback).  No compute calls: runs on the CPU-only box."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from supereight_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "se_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(se_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/se_b200.h but not exported"
    assert set(syms) == set(capi.EXPORTS), set(syms) ^ set(capi.EXPORTS)


def test_no_device_fails_loudly_not_silently():
    lib = capi.load_library()
    if lib.se_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.SeB200Error, match="no CUDA device"):
        capi.Map(capi.SE_B200_SDF, 256, 4.8, 160, 120)


def test_product_does_not_reference_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may touch oracle/; the fiber executor of the CPU test
    tier (tests/simt_emu) is likewise invisible to the package, to bench.py and to the driver's entry points."""
    forbidden = ("oracle_lib", "oracle/", "se_oracle", "liboracle", "seo_", "simt", "SIMT")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "supereight_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in forbidden:
                    assert pat not in text, f"{os.path.join(dirpath, f)} mentions {pat}"
    binary = open(capi.lib_path(), "rb").read()
    assert b"liboracle" not in binary and b"seo_" not in binary and b"simt" not in binary
    for f in ("bench.py", "__graft_entry__.py"):
        assert "simt" not in open(os.path.join(ROOT, f)).read(), f


def test_bspline_lut_matches_reference_table_checksum():
    """bfusion/bspline_lookup.cc:36-37: the product regenerates the table; its float32 bytes must hash to
    the checksum tests/golden/make_golden.py took from the reference's table."""
    lib = capi.load_library()
    lut = np.zeros(1000, np.float32)
    assert lib.se_b200_bspline_lut(lut.ctypes.data_as(C.c_void_p)) == 0
    want = open(os.path.join(ROOT, "tests", "golden", "bspline_lut.sha256")).read().split()[0]
    assert hashlib.sha256(lut.tobytes()).hexdigest() == want
    import oracle_lib
    o = np.array([oracle_lib.load().seo_bspline_lut(i) for i in range(1000)], np.float32)
    assert hashlib.sha256(o.tobytes()).hexdigest() == want
