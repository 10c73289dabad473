"""The on-disk protocol of a checkpoint (incubator_pegasus_b200/host/checkpoint_dir.h): write into checkpoint.<decree>.tmp, MANIFEST
last, rename; leftovers of interrupted attempts are cleared; manifests are parsed defensively.  C++ unit test
(tests/cpp/test_checkpoint_dir.cpp); the engine-side use is tests/test_checkpoint_gpu.py."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_directory_protocol(tmp_path):
    cxx = shutil.which("g++")
    assert cxx
    exe = str(tmp_path / "test_checkpoint_dir")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                           os.path.join(ROOT, "tests", "cpp", "test_checkpoint_dir.cpp"), "-o", exe])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    assert "runtime error" not in out.stderr, out.stderr
