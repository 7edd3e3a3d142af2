"""The C++ host facade (include/lapper_b200.hpp): compiles against the C ABI everywhere; on a GPU box
the reference's unit tests restated in C++ (tests/cpp/test_reference_suite.cpp) run through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "rust_lapper_b200")


def _build(tmp_path):
    exe = str(tmp_path / "test_reference_suite")
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "cpp", "test_reference_suite.cpp"),
        "-L", LIBDIR, "-llapper_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe,
    ])
    return exe


def test_cpp_facade_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "liblapper_b200.so")):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_reference_suite_runs(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all reference tests passed" in res.stdout
