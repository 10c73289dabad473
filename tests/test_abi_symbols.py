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


def test_router_needs_a_gpu_too(pgs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert pgs.lib().pgs_router_open(None, 0, C.byref(h)) == pgs.IO_ERROR and not h.value
    assert pgs.lib().pgs_router_device_for(None, 1, 3) == -1


def test_partition_index_is_the_reference_key_hash(pgs, oracle):
    # pegasus_key_hash % partition_count (pegasus_key_schema.h:150-165): the hash key decides, the sort key only when it is empty
    for hk, sk in ((b"user42", b"s"), (b"", b"only-sort"), (b"h", b"")):
        raw = len(hk).to_bytes(2, "big") + hk + sk
        want = oracle.lib().orc_key_hash(raw, len(raw)) % 256
        assert pgs.partition_index(hk, sk, 256) == want
    assert pgs.partition_index(b"a", b"x", 8) == pgs.partition_index(b"a", b"y", 8)


def test_product_does_not_link_the_oracle(pgs):
    import subprocess
    out = subprocess.run(["ldd", pgs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", pgs.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms
