"""The packed-sector arithmetic (rust_lapper_b200/csrc/packed_sector.cuh) compiled for the HOST and checked
against a brute-force Lapper::count -- the same header the CUDA kernels compile, so the format, the
non-power-of-two bucket division and the one-subtraction lane count are verified without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_sector_arithmetic_on_host(tmp_path):
    exe = str(tmp_path / "test_packed_sector")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "test_packed_sector.cpp"), "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "packed sector ok" in res.stdout
