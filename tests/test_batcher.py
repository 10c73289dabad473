"""The request-batching front end's host logic (incubator_pegasus_b200/host/batcher.h: the Coalescer's windows, the marshalling
into pgs_get_batch_multi-shaped calls, buffer growth, error fan-out) as a C++ unit test with a stand-in for the launch
(tests/cpp/test_batcher.cpp), built with ThreadSanitizer.  The device side of the same call is tests/test_zz_batcher_gpu.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_batcher_host_logic(tmp_path, sanitizer):
    cxx = shutil.which("g++")
    assert cxx, "g++ is part of this image"
    exe = str(tmp_path / "test_batcher")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-g", "-pthread", f"-fsanitize={sanitizer}",
                           os.path.join(ROOT, "tests", "cpp", "test_batcher.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    assert "WARNING: ThreadSanitizer" not in out.stderr and "runtime error" not in out.stderr, out.stderr
