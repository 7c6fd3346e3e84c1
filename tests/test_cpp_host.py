"""The C++ host mirror (include/poseidon252_b200.hpp) compiles against the C ABI and behaves like the
reference's API at the boundary.  CPU part: compile, link, host-only checks, loud failure without a GPU.
GPU part (-m gpu): the same binary exercises digest / encrypt / decrypt on the device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "poseidon252_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    from poseidon252_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", LIBDIR, "-lposeidon252_b200", "-Wl,-rpath," + LIBDIR])


def _run():
    return subprocess.run([EXE], capture_output=True, text=True, timeout=120)


def test_cpp_host_mirror_cpu():
    _build()
    res = _run()
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "host mirror ok" in res.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    _build()
    res = _run()
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "host mirror ok (GPU)" in res.stdout
