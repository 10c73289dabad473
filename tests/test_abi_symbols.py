"""CPU: libpegasus_b200.so loads without a GPU and exports every entry point include/pegasus_b200.h declares;
opening an engine without a CUDA device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    text = open(os.path.join(ROOT, "include", "pegasus_b200.h")).read()
    return sorted(set(re.findall(r"PGS_API[^;(]*?\b(pgs_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(pgs):
    L = pgs.lib()
    names = declared()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback(pgs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = pgs.lib().pgs_engine_open(None, C.byref(h))
    assert st == pgs.IO_ERROR and not h.value
    assert b"CUDA" in pgs.lib().pgs_last_error()


def test_product_does_not_link_the_oracle(pgs):
    import subprocess
    out = subprocess.run(["ldd", pgs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", pgs.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms
