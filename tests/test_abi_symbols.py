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


def test_header_is_plain_c_and_a_c_program_links(pgs, tmp_path):
    """the drop-in boundary is a C ABI: include/pegasus_b200.h compiles as strict C99 (no C++, no torch types) and a C program
    that uses nothing but the header links against the library and runs its host-side entry points"""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    assert cc
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "pegasus_b200.h"
int main(void)
{
    unsigned char key[64];
    const unsigned char hk[] = "hash", sk[] = "sort";
    int n = pgs_generate_key(hk, 4, sk, 4, key, sizeof key);
    if (n != 10 || key[0] != 0 || key[1] != 4 || memcmp(key + 2, "hashsort", 8)) return 1;
    if (pgs_partition_index(hk, 4, sk, 4, 8) >= 8) return 2;
    pgs_manual_compact_decision d;
    if (pgs_manual_compact_decide("manual_compact.disabled\0true\0", 1, 1000, 0, 0, 7, &d) != PGS_OK || !d.disabled) return 3;
    pgs_engine *e = NULL;
    if (pgs_engine_open(NULL, &e) == PGS_OK) pgs_engine_close(e); /* a GPU box: fine; here: it must fail, not fall back */
    printf("%d\n", (int)sizeof(pgs_get_result));
    return 0;
}
''')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(pgs.LIB_PATH)
    subprocess.check_call([cc, "-std=c99", "-pedantic-errors", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-lpegasus_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stderr)
    assert int(out.stdout.strip()) == C.sizeof(pgs.GetResult)
