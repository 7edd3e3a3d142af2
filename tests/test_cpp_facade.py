"""The C++ host facade (include/lapper_b200.hpp): compiles against the C ABI everywhere; on a GPU box
the reference's unit tests restated in C++ (tests/cpp/test_reference_suite.cpp) run through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "rust_lapper_b200")


def _build(tmp_path, name="test_reference_suite"):
    exe = str(tmp_path / name)
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
        "-L", LIBDIR, "-llapper_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe,
    ])
    return exe


def test_cpp_facade_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "liblapper_b200.so")):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(_build(tmp_path))
    assert os.path.exists(_build(tmp_path, "bincode_roundtrip"))


@pytest.mark.gpu
def test_cpp_reference_suite_runs(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all reference tests passed" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("coord", ["u32", "u64"])
def test_cpp_and_python_bincode_codecs_agree(tmp_path, coord):
    """`with_serde` (src/lib.rs:85-86, test :1387-1409): bytes written by the Python codec load through the C++ facade,
    which writes the same bytes back; a buffer the C++ side wrote loads in Python."""
    from rust_lapper_b200 import Lapper, Lapper64, bincode_io  # noqa: F401
    from test_bincode import SERDE_TEST_DATA

    exe = _build(tmp_path, "bincode_roundtrip")
    lapper = (Lapper if coord == "u32" else Lapper64)(SERDE_TEST_DATA)
    lapper.set_cov()
    raw = bincode_io.dumps(lapper, coord=coord, val="u32")
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(raw)
    res = subprocess.run([exe, coord, str(src), str(dst)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip() == "1"                                  # src/lib.rs:1407
    assert dst.read_bytes() == raw
    back = bincode_io.loads(dst.read_bytes(), coord=coord, val="u32", engine=coord)
    assert back.count(28974798, 33141355) == 1 and back.cov() == lapper.cov()
    src.write_bytes(raw[:-3])
    res = subprocess.run([exe, coord, str(src), str(dst)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "bincode" in res.stdout
