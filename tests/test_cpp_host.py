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


# ---- the Rust binding's extern "C" set, exercised from plain C ---------------------------------------------------
C_EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def _build_c():
    from poseidon252_b200 import build
    build.build()
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", C_EXE,
                           "-L", LIBDIR, "-lposeidon252_b200", "-Wl,-rpath," + LIBDIR])


def _c_calls(path):
    import re
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    return set(re.findall(r"\b(p252_[a-z0-9_]+)\s*\(", src))


def _rust_externs():
    """name -> number of parameters, parsed from the extern "C" block of bindings/rust/src/lib.rs"""
    import re
    src = open(os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    out = {}
    for name, params in re.findall(r"fn\s+(p252_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        out[name] = len([p for p in params.split(",") if p.strip()])
    return out


def _header_protos():
    import re
    src = open(os.path.join(ROOT, "include", "poseidon252_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(p252_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else len(params.split(","))
    return out


def test_rust_extern_block_matches_header_and_c_smoke():
    """The Rust source cannot be compiled here; instead: (1) every function its extern block declares exists in the
    header with the same number of parameters, and (2) tests/c/abi_smoke.c calls exactly that set."""
    rust, hdr = _rust_externs(), _header_protos()
    assert len(rust) >= 12
    for name, nparams in rust.items():
        assert name in hdr, name
        assert hdr[name] == nparams, (name, hdr[name], nparams)
    called = _c_calls(os.path.join(ROOT, "tests", "c", "abi_smoke.c"))
    assert called == set(rust), (sorted(called - set(rust)), sorted(set(rust) - called))


def test_c_abi_smoke_cpu():
    _build_c()
    res = subprocess.run([C_EXE], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "ABI_SMOKE_NO_DEVICE" in res.stdout or "ABI_SMOKE_OK" in res.stdout


@pytest.mark.gpu
def test_c_abi_smoke_gpu():
    _build_c()
    res = subprocess.run([C_EXE], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "ABI_SMOKE_OK" in res.stdout
