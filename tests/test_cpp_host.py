"""The C++ host mirror (include/islands_b200.hpp) compiles against the C ABI and behaves like the
reference interface; the GPU variant builds and searches an index through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_check")
    lib_dir = os.path.join(ROOT, "islands_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"), "-o", exe,
                           "-L", lib_dir, "-lislands_b200", f"-Wl,-rpath,{lib_dir}"])
    return exe


def test_cpp_host_mirror_cpu(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu(tmp_path, gpu_lib):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
