"""ctypes binding of the CPU oracle (oracle/lapper_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (rust_lapper_b200) never does.

`OracleLapper` mirrors the reference's `Lapper<u32, u32>` method for method (src/lib.rs:182-680)
so that tests read like the reference's own tests; intervals are (start, stop) tuples and `val`
is the interval's input position.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblapper_oracle.so")
_SO64 = os.path.join(_HERE, "liblapper_oracle64.so")  # the same source with I = u64 (-DLO_COORD_BITS=64)


def build(force: bool = False, coord_bits: int = 32) -> str:
    """Compile the oracle with gcc (plain C, no CUDA)."""
    so = _SO64 if coord_bits == 64 else _SO
    src = os.path.join(_HERE, "lapper_oracle.c")
    hdr = os.path.join(_HERE, "lapper_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr) if os.path.exists(f)
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.basename(so)] + (["-B"] if force else []))
    return so


def build_native(out_path: str) -> Optional[str]:
    """-march=native build into `out_path` (bench.py's CPU-baseline leg, on the GPU box's host)."""
    try:
        subprocess.check_call(["make", "-s", "-C", _HERE, "native", f"OUT={out_path}"])
        return out_path
    except Exception:
        return None


class _Binding:
    """ctypes view of one build of the oracle: the struct layouts and prototypes for I = u32 or I = u64"""

    def __init__(self, bits: int):
        self.bits = bits
        ct = C.c_uint64 if bits == 64 else C.c_uint32
        self.ct = ct
        self.np = np.uint64 if bits == 64 else np.uint32
        self.p = C.POINTER(ct)
        # lo_interval {I start; I stop; u32 val} (8-byte aligned, so padded to 24 bytes, when I = u64)
        self.iv_dtype = (np.dtype([("start", "<u8"), ("stop", "<u8"), ("val", "<u4"), ("pad", "<u4")]) if bits == 64 else
                         np.dtype([("start", "<u4"), ("stop", "<u4"), ("val", "<u4")]))

        class Iv(C.Structure):
            _fields_ = [("start", ct), ("stop", ct), ("val", C.c_uint32)]

        class Lapper(C.Structure):
            _fields_ = [
                ("intervals", C.POINTER(Iv)),
                ("starts", C.POINTER(ct)),
                ("stops", C.POINTER(ct)),
                ("len", C.c_size_t),
                ("cap", C.c_size_t),
                ("max_len", ct),
                ("cov_is_some", C.c_int),
                ("cov", ct),
                ("overlaps_merged", C.c_int),
            ]

        class IterFind(C.Structure):
            _fields_ = [("inner", C.POINTER(Lapper)), ("off", C.c_size_t), ("start", ct), ("stop", ct)]

        class IterDepth(C.Structure):
            _fields_ = [
                ("inner", C.POINTER(Lapper)),
                ("merged", C.POINTER(Lapper)),
                ("curr_merged_pos", ct),
                ("curr_pos", C.c_size_t),
                ("cursor", C.c_size_t),
                ("end", C.c_size_t),
            ]

        assert C.sizeof(Iv) == self.iv_dtype.itemsize
        self.Iv, self.Lapper, self.IterFind, self.IterDepth = Iv, Lapper, IterFind, IterDepth
        self.lib = None

    def bind(self, lib):
        ct, cp = self.ct, self.p
        LP = C.POINTER(self.Lapper)
        lib.lo_new.restype = LP
        lib.lo_new.argtypes = [cp, cp, _u32p, C.c_size_t]
        lib.lo_free.argtypes = [LP]
        lib.lo_insert.argtypes = [LP, ct, ct, C.c_uint32]
        lib.lo_interval_intersect.restype = ct
        lib.lo_interval_intersect.argtypes = [ct] * 4
        lib.lo_lower_bound.restype = C.c_size_t
        lib.lo_lower_bound.argtypes = [ct, C.POINTER(self.Iv), C.c_size_t]
        lib.lo_bsearch_seq.restype = C.c_size_t
        lib.lo_bsearch_seq.argtypes = [ct, cp, C.c_size_t]
        lib.lo_count.restype = C.c_uint64
        lib.lo_count.argtypes = [LP, ct, ct]
        lib.lo_find.restype = self.IterFind
        lib.lo_find.argtypes = [LP, ct, ct]
        lib.lo_seek.restype = self.IterFind
        lib.lo_seek.argtypes = [LP, ct, ct, C.POINTER(C.c_size_t)]
        lib.lo_iter_find_next.restype = C.c_int64
        lib.lo_iter_find_next.argtypes = [C.POINTER(self.IterFind)]
        for f in ("lo_calculate_coverage", "lo_cov", "lo_set_cov"):
            getattr(lib, f).restype = ct
            getattr(lib, f).argtypes = [LP]
        lib.lo_merge_overlaps.argtypes = [LP]
        lib.lo_union_and_intersect.argtypes = [LP, LP, cp, cp]
        lib.lo_depth.restype = self.IterDepth
        lib.lo_depth.argtypes = [LP]
        lib.lo_iter_depth_next.restype = C.c_int
        lib.lo_iter_depth_next.argtypes = [C.POINTER(self.IterDepth), cp, cp, _u32p]
        lib.lo_iter_depth_free.argtypes = [C.POINTER(self.IterDepth)]
        lib.lo_count_batch.argtypes = [LP, cp, cp, C.c_size_t, _u64p, C.c_int]
        lib.lo_find_batch_offsets.restype = C.c_uint64
        lib.lo_find_batch_offsets.argtypes = [LP, cp, cp, C.c_size_t, _u64p, C.c_int]
        lib.lo_find_batch_fill.argtypes = [LP, cp, cp, C.c_size_t, _u64p, _u32p, C.c_int]
        lib.lo_seek_batch_fill.argtypes = [LP, cp, cp, C.c_size_t, _u64p, _u32p]
        lib.lo_find_count_sum.restype = C.c_uint64
        lib.lo_find_count_sum.argtypes = [LP, cp, cp, C.c_size_t, C.c_int]
        lib._binding = self
        return lib

    def arr(self, a) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(a, dtype=self.np))

    def ptr(self, a: np.ndarray):
        return a.ctypes.data_as(self.p)


_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_bindings = {32: _Binding(32), 64: _Binding(64)}
_IterFind = _bindings[32].IterFind


def load(path: Optional[str] = None, coord_bits: int = 32):
    """The oracle library for I = u32 (default) or I = u64; `path` loads another build of the u32 one."""
    b = _bindings[coord_bits]
    if path is None:
        if b.lib is None:
            b.lib = b.bind(C.CDLL(build(coord_bits=coord_bits)))
        return b.lib
    return _Binding(coord_bits).bind(C.CDLL(path))


def _as_u32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


def _p32(a: np.ndarray):
    return a.ctypes.data_as(_u32p)


def _p64(a: np.ndarray):
    return a.ctypes.data_as(_u64p)


class OracleLapper:
    """`Lapper<u32,u32>` (coord_bits=64: `Lapper<u64,u32>`) of the reference, restated on the CPU (src/lib.rs:182-680)."""

    def __init__(self, intervals: Iterable[Tuple[int, int]] = (), *, starts=None, stops=None, vals=None, lib=None,
                 coord_bits: int = 32):
        self._lib = lib or load(coord_bits=coord_bits)
        self._b = self._lib._binding
        if starts is None:
            ivs = list(intervals)
            starts = [iv[0] for iv in ivs]
            stops = [iv[1] for iv in ivs]
            if ivs and len(ivs[0]) > 2:
                vals = [iv[2] for iv in ivs]
        s, e = self._b.arr(starts), self._b.arr(stops)
        assert s.shape == e.shape
        v = _as_u32(vals) if vals is not None else None
        self._h = self._lib.lo_new(self._b.ptr(s), self._b.ptr(e), _p32(v) if v is not None else None, s.size)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lo_free(h)

    # -- structure (private fields the reference's insert_equality_* tests compare, src/lib.rs:907-1019)
    def __len__(self) -> int:
        return int(self._h.contents.len)

    def is_empty(self) -> bool:
        return len(self) == 0

    @property
    def max_len(self) -> int:
        return int(self._h.contents.max_len)

    @property
    def overlaps_merged(self) -> bool:
        return bool(self._h.contents.overlaps_merged)

    def _arr(self, field: str) -> np.ndarray:
        n = len(self)
        if n == 0:
            return np.zeros(0, self._b.np)
        return np.ctypeslib.as_array(getattr(self._h.contents, field), shape=(n,)).copy()

    @property
    def starts(self) -> np.ndarray:
        return self._arr("starts")

    @property
    def stops(self) -> np.ndarray:
        return self._arr("stops")

    def intervals_soa(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        n = len(self)
        if n == 0:
            return np.zeros(0, self._b.np), np.zeros(0, self._b.np), np.zeros(0, np.uint32)
        nbytes = n * self._b.iv_dtype.itemsize
        raw = np.frombuffer(C.string_at(self._h.contents.intervals, nbytes), dtype=self._b.iv_dtype)
        return raw["start"].copy(), raw["stop"].copy(), raw["val"].copy()

    @property
    def intervals(self) -> List[Tuple[int, int]]:
        s, e, _ = self.intervals_soa()
        return list(zip(s.tolist(), e.tolist()))

    def iter(self):
        return iter(self.intervals)

    def insert(self, start: int, stop: int, val: int = 0) -> None:
        self._lib.lo_insert(self._h, start, stop, val)

    # -- queries
    def count(self, start: int, stop: int) -> int:
        return int(self._lib.lo_count(self._h, start, stop))

    def _drain(self, it) -> List[int]:
        out = []
        while True:
            p = self._lib.lo_iter_find_next(C.byref(it))
            if p < 0:
                return out
            out.append(int(p))

    def find_idx(self, start: int, stop: int) -> List[int]:
        return self._drain(self._lib.lo_find(self._h, start, stop))

    def find(self, start: int, stop: int) -> List[Tuple[int, int]]:
        ivs = self.intervals
        return [ivs[i] for i in self.find_idx(start, stop)]

    def seek_idx(self, start: int, stop: int, cursor: List[int]) -> List[int]:
        """cursor is a one-element list standing in for `&mut usize`."""
        c = C.c_size_t(cursor[0])
        it = self._lib.lo_seek(self._h, start, stop, C.byref(c))
        cursor[0] = int(c.value)
        return self._drain(it)

    def seek(self, start: int, stop: int, cursor: List[int]) -> List[Tuple[int, int]]:
        ivs = self.intervals
        return [ivs[i] for i in self.seek_idx(start, stop, cursor)]

    # -- set ops
    def cov(self) -> int:
        return int(self._lib.lo_cov(self._h))

    def set_cov(self) -> int:
        return int(self._lib.lo_set_cov(self._h))

    def merge_overlaps(self) -> None:
        self._lib.lo_merge_overlaps(self._h)

    def union_and_intersect(self, other: "OracleLapper") -> Tuple[int, int]:
        u, i = self._b.ct(0), self._b.ct(0)
        self._lib.lo_union_and_intersect(self._h, other._h, C.byref(u), C.byref(i))
        return int(u.value), int(i.value)

    def intersect(self, other: "OracleLapper") -> int:
        return self.union_and_intersect(other)[1]

    def union(self, other: "OracleLapper") -> int:
        return self.union_and_intersect(other)[0]

    def depth(self) -> List[Tuple[int, int, int]]:
        if len(self) == 0:
            raise IndexError("depth() on an empty lapper panics in the reference (src/lib.rs:744)")
        it = self._lib.lo_depth(self._h)
        out = []
        s, e, v = self._b.ct(0), self._b.ct(0), C.c_uint32(0)
        while self._lib.lo_iter_depth_next(C.byref(it), C.byref(s), C.byref(e), C.byref(v)):
            out.append((int(s.value), int(e.value), int(v.value)))
        self._lib.lo_iter_depth_free(C.byref(it))
        return out

    # -- batch drivers
    def count_batch(self, q_start, q_stop, nthreads: int = 1) -> np.ndarray:
        qs, qe = self._b.arr(q_start), self._b.arr(q_stop)
        out = np.empty(qs.size, np.uint64)
        self._lib.lo_count_batch(self._h, self._b.ptr(qs), self._b.ptr(qe), qs.size, _p64(out), nthreads)
        return out

    def find_batch(self, q_start, q_stop, nthreads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        qs, qe = self._b.arr(q_start), self._b.arr(q_stop)
        offsets = np.empty(qs.size + 1, np.uint64)
        total = int(self._lib.lo_find_batch_offsets(self._h, self._b.ptr(qs), self._b.ptr(qe), qs.size, _p64(offsets), nthreads))
        indices = np.empty(total, np.uint32)
        self._lib.lo_find_batch_fill(self._h, self._b.ptr(qs), self._b.ptr(qe), qs.size, _p64(offsets), _p32(indices), nthreads)
        return offsets, indices

    def seek_batch(self, q_start, q_stop, offsets: np.ndarray) -> np.ndarray:
        qs, qe = self._b.arr(q_start), self._b.arr(q_stop)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        indices = np.full(int(offsets[-1]), 0xFFFFFFFF, np.uint32)
        self._lib.lo_seek_batch_fill(self._h, self._b.ptr(qs), self._b.ptr(qe), qs.size, _p64(offsets), _p32(indices))
        return indices

    def find_count_sum(self, q_start, q_stop, nthreads: int = 1) -> int:
        qs, qe = self._b.arr(q_start), self._b.arr(q_stop)
        return int(self._lib.lo_find_count_sum(self._h, self._b.ptr(qs), self._b.ptr(qe), qs.size, nthreads))


def interval_intersect(a: Sequence[int], b: Sequence[int]) -> int:
    """Interval::intersect (src/lib.rs:130-136)."""
    return int(load().lo_interval_intersect(a[0], a[1], b[0], b[1]))


# ---------------------------------------------------------------- independent set definition
def brute_find(starts, stops, a: int, b: int) -> List[int]:
    """Ten-line set definition: ascending sorted positions i with
    start_i < b and stop_i > a, over the stable (start, stop)-sorted order.  Independent of the
    oracle's search loops; used to cross-check the oracle itself."""
    s = np.asarray(starts, dtype=np.int64)
    e = np.asarray(stops, dtype=np.int64)
    order = np.lexsort((e, s))  # stable, primary key start, secondary stop
    ss, ee = s[order], e[order]
    return np.nonzero((ss < b) & (ee > a))[0].tolist()
