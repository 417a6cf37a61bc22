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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/islands_b200.h compiles as C99 with -pedantic (the form cgo / bindgen / ctypes consumers bind) and a C
    program links against the library and gets the reference's defaults and errors back (tests/c/abi_check.c)."""
    exe = str(tmp_path / "abi_check")
    lib_dir = os.path.join(ROOT, "islands_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-o", exe, "-L", lib_dir, "-lislands_b200",
                           f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
