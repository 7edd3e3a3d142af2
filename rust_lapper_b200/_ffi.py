"""ctypes binding of liblapper_b200.so (the C ABI of include/lapper_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
fallback of any kind: if the library is missing, or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LAPPER_B200_LIB selects an alternative build of the same library (kernel A/B experiments only)
LIB_PATH = os.environ.get("LAPPER_B200_LIB") or os.path.join(_HERE, "liblapper_b200.so")

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_OOM, ERR_PRECONDITION, ERR_NO_DEVICE, ERR_CAPACITY = 1, 2, 3, 4, 5, 6

BUILD_DEFAULT = 0
BUILD_NO_BUCKETS = 1
BUILD_FORCE_BUCKETS = 2
BUILD_NO_DENSE = 4
BUILD_FORCE_DENSE = 8


def build_shift(k: int) -> int:
    return (k + 1) << 8


def build_dense_width(w: int) -> int:
    return w << 16


class LapperError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liblapper_b200 error {code}: {msg}")
        self.code = code


class DevView(C.Structure):
    _fields_ = [
        ("starts", C.c_void_p),
        ("stops_by_iv", C.c_void_p),
        ("stops_sorted", C.c_void_p),
        ("perm", C.c_void_p),
        ("prefix_max", C.c_void_p),
        ("n", C.c_uint64),
    ]


_vp, _u32, _u64, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/lapper_b200.h declares
SIGNATURES = {
    "lapper_version": (_int, []),
    "lapper_last_error": (C.c_char_p, []),
    "lapper_device_count": (_int, [C.POINTER(_int)]),
    "lapper_query_kernel_launches": (_u64, []),
    "lapper_debug_violations": (C.c_longlong, []),
    "lapper_debug_set_sorted_min_queries": (_u64, [_u64]),
    "lapper_debug_sorted_run_stats": (_int, [_vp, _int]),
    "lapper_build_u32": (_int, [_vp, _vp, _u64, _int, _pvp]),
    "lapper_build_u32_ex": (_int, [_vp, _vp, _u64, _int, _u32, _pvp]),
    "lapper_build_u32_dev": (_int, [_vp, _vp, _u64, _int, _u32, _vp, _pvp]),
    "lapper_free": (None, [_vp]),
    "lapper_len": (_u64, [_vp]),
    "lapper_max_len": (_u32, [_vp]),
    "lapper_overlaps_merged": (_int, [_vp]),
    "lapper_has_inverted": (_int, [_vp]),
    "lapper_device": (_int, [_vp]),
    "lapper_bucket_shift": (_u32, [_vp]),
    "lapper_index_bytes": (_u64, [_vp]),
    "lapper_packed_table_info": (_int, [_vp, C.POINTER(_u32), C.POINTER(_u64), C.POINTER(_u64)]),
    "lapper_tune_query_length": (_int, [_vp, _u32]),
    "lapper_get_intervals": (_int, [_vp, _vp, _vp, _vp]),
    "lapper_get_starts": (_int, [_vp, _vp]),
    "lapper_get_stops": (_int, [_vp, _vp]),
    "lapper_get_dev_view": (_int, [_vp, C.POINTER(DevView)]),
    "lapper_count_batch_u32": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "lapper_count_batch_u32_c32": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "lapper_count_batch_u32_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "lapper_count_batch_u32_dev64": (_int, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "lapper_find_batch_u32": (_int, [_vp, _vp, _vp, _u64, _pvp]),
    "lapper_find_batch_u32_devq": (_int, [_vp, _vp, _vp, _u64, _vp, _pvp]),
    "lapper_result_nq": (_u64, [_vp]),
    "lapper_result_total": (_u64, [_vp]),
    "lapper_result_copy": (_int, [_vp, _vp, _vp]),
    "lapper_result_dev": (_int, [_vp, _pvp, _pvp]),
    "lapper_result_free": (None, [_vp]),
    "lapper_find_batch_u32_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64), _vp]),
    "lapper_find_batch_offsets_u32_dev": (_int, [_vp, _vp, _vp, _u64, _vp, C.POINTER(_u64), _vp]),
    "lapper_find_batch_fill_u32_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "lapper_find_batch_fill_total_u32_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "lapper_find_batch_u32_host": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64)]),
    "lapper_seek_cursor": (_int, [_vp, _u32, C.POINTER(_u64)]),
    "lapper_get_cached_state": (_int, [_vp, C.POINTER(_int), C.POINTER(_u32), C.POINTER(_int)]),
    "lapper_restore_cached_state": (_int, [_vp, _int, _u32, _int]),
    "lapper_insert_u32": (_int, [_vp, _u32, _u32, C.POINTER(_u64)]),
    "lapper_merge_overlaps": (_int, [_vp]),
    "lapper_cov": (_int, [_vp, C.POINTER(_u32)]),
    "lapper_set_cov": (_int, [_vp, C.POINTER(_u32)]),
    "lapper_union_and_intersect": (_int, [_vp, _vp, C.POINTER(_u32), C.POINTER(_u32)]),
    "lapper_depth": (_int, [_vp, C.POINTER(_u64), _pvp, _pvp, _pvp]),
    "lapper_host_free": (None, [_vp]),
    "lapper_replicate": (_int, [_vp, C.POINTER(_int), _int, _pvp]),
    "lapper_group_free": (None, [_vp]),
    "lapper_group_size": (_int, [_vp]),
    "lapper_group_replica": (_vp, [_vp, _int]),
    "lapper_group_count_batch_u32": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "lapper_group_find_batch_u32": (_int, [_vp, _vp, _vp, _u64, _vp, _pvp, C.POINTER(_u64)]),
    "lapper_group_cov": (_int, [_vp, C.POINTER(_u32)]),
    "lapper_group_set_cov": (_int, [_vp, C.POINTER(_u32)]),
    "lapper_group_merge_overlaps": (_int, [_vp]),
    "lapper_from_sorted_dev": (_int, [_vp, _vp, _vp, _vp, _u64, _int, _int, _u32, _vp, _pvp]),
    "lapper_host_alloc_pinned": (_int, [_u64, _pvp]),
    # I = u64 (csrc/engine64.cu)
    "lapper64_build": (_int, [_vp, _vp, _u64, _int, _pvp]),
    "lapper64_build_dev": (_int, [_vp, _vp, _u64, _int, _vp, _pvp]),
    "lapper64_free": (None, [_vp]),
    "lapper64_len": (_u64, [_vp]),
    "lapper64_max_len": (_u64, [_vp]),
    "lapper64_overlaps_merged": (_int, [_vp]),
    "lapper64_has_inverted": (_int, [_vp]),
    "lapper64_get_intervals": (_int, [_vp, _vp, _vp, _vp]),
    "lapper64_get_sorted_stops": (_int, [_vp, _vp]),
    "lapper64_count_batch": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "lapper64_count_batch_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "lapper64_find_batch": (_int, [_vp, _vp, _vp, _u64, _pvp]),
    "lapper64_find_batch_dev": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64), _vp]),
    "lapper64_cov": (_int, [_vp, C.POINTER(_u64)]),
    "lapper64_set_cov": (_int, [_vp, C.POINTER(_u64)]),
    "lapper64_get_cached_state": (_int, [_vp, C.POINTER(_int), C.POINTER(_u64), C.POINTER(_int)]),
    "lapper64_restore_cached_state": (_int, [_vp, _int, _u64, _int]),
    "lapper64_merge_overlaps": (_int, [_vp]),
    "lapper64_union_and_intersect": (_int, [_vp, _vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "lapper64_insert": (_int, [_vp, _u64, _u64, C.POINTER(_u64)]),
    "lapper64_seek_cursor": (_int, [_vp, _u64, C.POINTER(_u64)]),
    "lapper64_depth": (_int, [_vp, C.POINTER(_u64), _pvp, _pvp, _pvp]),
    "lapper64_is_narrow": (_int, [_vp]),
    "lapper64_tune_query_length": (_int, [_vp, _u32]),
    "lapper64_debug_set_narrow_mode": (_int, [_int]),
    "lapper_host_free_pinned": (None, [_vp]),
}

_lib = None


def lib():
    """Load liblapper_b200.so.  Raises (never falls back) when the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  rust_lapper_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().lapper_last_error()
        raise LapperError(rc, msg.decode("utf-8", "replace") if msg else "")
