"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lapper_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lapper_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lapper(?:64)?_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from rust_lapper_b200 import _ffi

    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _ffi.lib()


def test_header_declares_the_reference_api():
    syms = declared_symbols()
    for must in ("lapper_build_u32", "lapper_count_batch_u32", "lapper_find_batch_u32", "lapper_merge_overlaps",
                 "lapper_cov", "lapper_set_cov", "lapper_union_and_intersect", "lapper_seek_cursor", "lapper_free",
                 "lapper_find_batch_u32_dev", "lapper64_build", "lapper64_count_batch", "lapper64_find_batch",
                 "lapper64_merge_overlaps", "lapper64_cov", "lapper64_union_and_intersect"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in lapper_b200.h but not exported: {missing}"


def test_ffi_table_matches_header(lib):
    from rust_lapper_b200 import _ffi

    assert sorted(_ffi.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_device(lib):
    count = C.c_int(-1)
    assert lib.lapper_device_count(C.byref(count)) == 0
    if count.value > 0:
        pytest.skip("a CUDA device is present")
    from rust_lapper_b200 import Lapper, LapperError, _ffi

    with pytest.raises(LapperError) as ei:
        Lapper([(1, 5), (3, 9)])
    assert ei.value.code == _ffi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.lapper_last_error()


def test_null_arguments_are_status_codes_not_crashes(lib):
    from rust_lapper_b200 import _ffi

    h = C.c_void_p(0)
    assert lib.lapper_build_u32(None, None, 5, 0, C.byref(h)) in (_ffi.ERR_INVALID_ARG, _ffi.ERR_NO_DEVICE)
    assert lib.lapper_count_batch_u32(None, None, None, 3, None) == _ffi.ERR_INVALID_ARG
    assert lib.lapper_len(None) == 0
    lib.lapper_free(None)
    lib.lapper_result_free(None)
    lib.lapper_group_free(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rust_lapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "lapper_oracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)


def test_synth_host_and_torch_agree():
    from rust_lapper_b200 import synth

    for fn, args in [
        (synth.intervals_uniform, (42, 5000, synth.GRCH38, 50, 5000)),
        (synth.intervals_gene_exon, (44, 5000)),
        (synth.intervals_pathological, (46, 5000)),
        (synth.queries_fixed_len, (43, 5000, synth.GRCH38, 150)),
        (synth.queries_sorted_fixed_len, (45, 5000, synth.GRCH38, 150)),
    ]:
        a = [synth.to_numpy_u32(x) for x in fn(*args)]
        b = [synth.to_numpy_u32(x) for x in fn(*args, device="cpu")]
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), fn.__name__
        assert np.all(a[1].astype(np.int64) >= a[0].astype(np.int64))
    s, _ = synth.queries_sorted_fixed_len(45, 5000, synth.GRCH38, 150)
    assert np.all(np.diff(s.astype(np.int64)) >= 0)


def test_python_mirror_rejects_out_of_range_coordinates():
    """int64 / negative inputs used to wrap modulo 2^32 silently."""
    import numpy as np
    import pytest

    from rust_lapper_b200.lapper import _u32

    assert _u32([0, 4294967295]).dtype == np.uint32
    assert _u32(np.array([5, 7], dtype=np.int64)).tolist() == [5, 7]
    assert _u32(np.array([], dtype=np.int64)).size == 0
    for bad in ([-1, 3], [1 << 32], np.array([2**40], dtype=np.int64), [1.5]):
        with pytest.raises(ValueError):
            _u32(bad)


def test_header_is_plain_c_and_links_from_a_c_program(lib, tmp_path):
    """The boundary is a C ABI: include/lapper_b200.h compiles as strict C99 (no C++ in the signatures) and a C program
    links against the library and gets status codes, not exceptions (here: no device / null arguments)."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "lapper_b200.h"\n'
        "int main(void) {\n"
        "    lapper_t *h = 0;\n"
        "    lapper64_t *w = 0;\n"
        "    uint32_t s[2] = {1, 5}, e[2] = {4, 9};\n"
        "    uint64_t s64[2] = {1, 5}, e64[2] = {4, 9};\n"
        "    int devices = -1;\n"
        "    if (lapper_version() <= 0 || lapper_device_count(&devices) != LAPPER_OK) return 1;\n"
        "    if (lapper_build_u32(s, e, 2, 0, 0) == LAPPER_OK) return 2;         /* null output handle */\n"
        "    if (devices == 0 && lapper_build_u32(s, e, 2, 0, &h) != LAPPER_ERR_NO_DEVICE) return 3;\n"
        "    if (devices == 0 && lapper64_build(s64, e64, 2, 0, &w) != LAPPER_ERR_NO_DEVICE) return 4;\n"
        '    printf("%d %s\\n", devices, lapper_last_error());\n'
        "    lapper_free(h);\n"
        "    lapper64_free(w);\n"
        "    return 0;\n"
        "}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "rust_lapper_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-L", libdir, "-llapper_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
