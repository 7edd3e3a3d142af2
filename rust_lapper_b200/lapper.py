"""Host-side mirror of rust-lapper's `Lapper<u32, T>` (reference: /root/reference/src/lib.rs:182-680).

Same method names and argument meaning as the reference -- new / find / seek / count /
merge_overlaps / cov / set_cov / union_and_intersect / intersect / union -- plus the batch entry
points `count_batch` and `find_batch` (CSR offsets + positions).  Every method calls the CUDA
engine through the C ABI; `val`s stay on the host and are looked up through `perm`.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Iterable, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _ffi


class Interval(NamedTuple):
    """`Interval<I, T>` (src/lib.rs:88-100): half-open [start, stop) with a payload."""

    start: int
    stop: int
    val: Any = None

    def intersect(self, other: "Interval") -> int:
        """src/lib.rs:130-136: min(stops).checked_sub(max(starts)).unwrap_or(0)"""
        lo, hi = max(self.start, other.start), min(self.stop, other.stop)
        return hi - lo if hi >= lo else 0

    def overlap(self, start: int, stop: int) -> bool:
        """src/lib.rs:138-142"""
        return self.start < stop and self.stop > start


def _u32(a) -> np.ndarray:
    """Coordinates as a contiguous u32 array.  Values outside [0, 2^32) are an error, never wrapped: the engine is
    I = u32 (the reference would not compile `Interval<u32, _>` with such a literal either)."""
    arr = np.asarray(a)
    if arr.dtype != np.uint32:
        if arr.size and arr.dtype.kind not in "iub":
            if arr.dtype.kind == "f" and np.all(np.isfinite(arr)) and np.all(arr == np.floor(arr)):
                arr = arr.astype(np.float64)
            else:
                raise ValueError(f"coordinates must be integers, got dtype {arr.dtype}")
        if arr.size and (arr.min() < 0 or arr.max() > 0xFFFFFFFF):
            raise ValueError("coordinate out of range for u32 (0 <= x <= 4294967295)")
        arr = arr.astype(np.uint32)
    return np.ascontiguousarray(arr).reshape(-1)


def _ptr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class Lapper:
    """`Lapper<u32, T>` on one B200.  `intervals` are (start, stop[, val]) tuples or, with
    `starts=`/`stops=`, two integer arrays (vals optional)."""

    def __init__(self, intervals: Iterable[Sequence] = (), *, starts=None, stops=None, vals=None,
                 device: int = 0, flags: int = _ffi.BUILD_DEFAULT):
        self._lib = _ffi.lib()
        self._h = C.c_void_p(0)
        if starts is None:
            ivs = list(intervals)
            starts = [iv[0] for iv in ivs]
            stops = [iv[1] for iv in ivs]
            if ivs and len(ivs[0]) > 2:
                vals = [iv[2] for iv in ivs]
        s, e = _u32(starts), _u32(stops)
        if s.shape != e.shape:
            raise ValueError("starts and stops differ in length")
        self._vals = list(vals) if vals is not None else None
        self._cache = None
        _ffi.check(self._lib.lapper_build_u32_ex(_ptr(s), _ptr(e), s.size, device, flags, C.byref(self._h)))

    @classmethod
    def _from_handle(cls, handle, vals=None):
        self = cls.__new__(cls)
        self._lib = _ffi.lib()
        self._h = handle
        self._vals = vals
        self._cache = None
        return self

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), C.c_void_p(0)
        if h:
            self._lib.lapper_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- structure
    @property
    def handle(self):
        return self._h

    def __len__(self) -> int:  # src/lib.rs:284-287
        return int(self._lib.lapper_len(self._h))

    def len(self) -> int:
        return len(self)

    def is_empty(self) -> bool:  # src/lib.rs:296-299
        return len(self) == 0

    @property
    def max_len(self) -> int:
        return int(self._lib.lapper_max_len(self._h))

    @property
    def overlaps_merged(self) -> bool:
        return bool(self._lib.lapper_overlaps_merged(self._h))

    @property
    def has_inverted(self) -> bool:
        """some interval has stop < start (count is defined for such sets, src/lib.rs:610-618; cov / union are not)"""
        return bool(self._lib.lapper_has_inverted(self._h))

    @property
    def bucket_shift(self) -> Optional[int]:
        k = int(self._lib.lapper_bucket_shift(self._h))
        return None if k == 0xFFFFFFFF else k

    def tune_query_length(self, typical_query_len: int) -> None:
        """Announce the typical query length (stop - start): the packed count table is rebuilt with a margin that covers
        it (faster random-order `count` for queries longer than 150-bp reads; results never change)."""
        _ffi.check(self._lib.lapper_tune_query_length(self._h, int(typical_query_len)))

    @property
    def packed_table(self) -> Optional[Tuple[int, int, int]]:
        """(bucket width, sectors, overflowed sectors) of the packed table, None when it was not built."""
        w, ns, no = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        _ffi.check(self._lib.lapper_packed_table_info(self._h, C.byref(w), C.byref(ns), C.byref(no)))
        return (int(w.value), int(ns.value), int(no.value)) if w.value else None

    def _sorted(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if self._cache is None:
            n = len(self)
            s, e, p = (np.empty(n, np.uint32) for _ in range(3))
            _ffi.check(self._lib.lapper_get_intervals(self._h, _ptr(s), _ptr(e), _ptr(p)))
            self._cache = (s, e, p)
        return self._cache

    def intervals_soa(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(starts, stops, perm) of the sorted interval vector; perm = input position."""
        return self._sorted()

    @property
    def intervals(self) -> List[Tuple[int, int]]:
        s, e, _ = self._sorted()
        return list(zip(s.tolist(), e.tolist()))

    def interval_at(self, pos: int) -> Interval:
        s, e, p = self._sorted()
        val = self._vals[int(p[pos])] if self._vals is not None else None
        return Interval(int(s[pos]), int(e[pos]), val)

    def iter(self):  # src/lib.rs:353-359
        return iter(self.intervals)

    def __iter__(self):
        return self.iter()

    @property
    def starts(self) -> np.ndarray:  # private `starts` (src/lib.rs:114)
        out = np.empty(len(self), np.uint32)
        _ffi.check(self._lib.lapper_get_starts(self._h, _ptr(out)))
        return out

    @property
    def stops(self) -> np.ndarray:  # private sorted `stops` (src/lib.rs:116)
        out = np.empty(len(self), np.uint32)
        _ffi.check(self._lib.lapper_get_stops(self._h, _ptr(out)))
        return out

    # ---------------------------------------------------------------- batch queries
    def count_batch(self, q_start, q_stop) -> np.ndarray:
        """counts[j] == count(q_start[j], q_stop[j]) as the reference's usize (u64)."""
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        out = np.empty(qs.size, np.uint64)
        _ffi.check(self._lib.lapper_count_batch_u32(self._h, _ptr(qs), _ptr(qe), qs.size, _ptr(out)))
        return out

    def count_batch_c32(self, q_start, q_stop) -> np.ndarray:
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        out = np.empty(qs.size, np.uint32)
        _ffi.check(self._lib.lapper_count_batch_u32_c32(self._h, _ptr(qs), _ptr(qe), qs.size, _ptr(out)))
        return out

    def find_batch(self, q_start, q_stop) -> Tuple[np.ndarray, np.ndarray]:
        """CSR: indices[offsets[j]:offsets[j+1]] are the positions find(q_j) yields, in order."""
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        res = C.c_void_p(0)
        _ffi.check(self._lib.lapper_find_batch_u32(self._h, _ptr(qs), _ptr(qe), qs.size, C.byref(res)))
        try:
            total = int(self._lib.lapper_result_total(res))
            offsets = np.empty(qs.size + 1, np.uint64)
            indices = np.empty(total, np.uint32)
            _ffi.check(self._lib.lapper_result_copy(res, _ptr(offsets), _ptr(indices)))
        finally:
            self._lib.lapper_result_free(res)
        return offsets, indices

    def find_batch_into(self, q_start, q_stop, offsets_out: np.ndarray, indices_out: Optional[np.ndarray]) -> int:
        """find_batch straight into caller-owned host buffers (pinned ones give full PCIe speed) through the pipelined
        one-call entry point; returns the total.  indices_out=None asks for offsets and the total only; a buffer that
        is too small raises LapperError(ERR_CAPACITY) after offsets_out has been filled."""
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        if offsets_out.dtype != np.uint64 or offsets_out.size < qs.size + 1 or not offsets_out.flags.c_contiguous:
            raise ValueError("offsets_out must be a contiguous uint64 array of nq + 1 entries")
        if indices_out is not None and (indices_out.dtype != np.uint32 or not indices_out.flags.c_contiguous):
            raise ValueError("indices_out must be a contiguous uint32 array")
        total = C.c_uint64(0)
        cap = 0 if indices_out is None else indices_out.size
        ip = C.c_void_p(indices_out.ctypes.data) if indices_out is not None and indices_out.size else C.c_void_p(0)
        _ffi.check(self._lib.lapper_find_batch_u32_host(self._h, _ptr(qs), _ptr(qe), qs.size, C.c_void_p(offsets_out.ctypes.data),
                                                        ip, cap, C.byref(total)))
        return int(total.value)

    # ---------------------------------------------------------------- per-query API of the reference
    def count(self, start: int, stop: int) -> int:  # src/lib.rs:610-618
        return int(self.count_batch([start], [stop])[0])

    def find_idx(self, start: int, stop: int) -> List[int]:
        _, idx = self.find_batch([start], [stop])
        return idx.tolist()

    def find(self, start: int, stop: int) -> List[Tuple[int, int]]:  # src/lib.rs:628-639
        s, e, _ = self._sorted()
        return [(int(s[i]), int(e[i])) for i in self.find_idx(start, stop)]

    def seek_idx(self, start: int, stop: int, cursor: List[int]) -> List[int]:
        """src/lib.rs:656-679.  `cursor` is a one-element list standing in for `&mut usize`; it is
        updated exactly as the reference updates it and the scan starts there."""
        c = C.c_uint64(cursor[0])
        _ffi.check(self._lib.lapper_seek_cursor(self._h, start, C.byref(c)))
        cursor[0] = int(c.value)
        return [i for i in self.find_idx(start, stop) if i >= cursor[0]]

    def seek(self, start: int, stop: int, cursor: List[int]) -> List[Tuple[int, int]]:
        s, e, _ = self._sorted()
        return [(int(s[i]), int(e[i])) for i in self.seek_idx(start, stop, cursor)]

    def insert(self, start: int, stop: int, val: Any = None) -> None:
        """src/lib.rs:260-273: sorted insert (before equal intervals); clears cov and overlaps_merged."""
        pos = C.c_uint64(0)
        _ffi.check(self._lib.lapper_insert_u32(self._h, start, stop, C.byref(pos)))
        if self._vals is not None:
            while len(self._vals) < int(pos.value):
                self._vals.append(None)
            self._vals.append(val)
        self._cache = None

    # ---------------------------------------------------------------- set operations
    def merge_overlaps(self) -> None:  # src/lib.rs:361-416
        _ffi.check(self._lib.lapper_merge_overlaps(self._h))
        self._cache = None

    def cov(self) -> int:  # src/lib.rs:310-316
        out = C.c_uint32(0)
        _ffi.check(self._lib.lapper_cov(self._h, C.byref(out)))
        return int(out.value)

    def set_cov(self) -> int:  # src/lib.rs:320-324
        out = C.c_uint32(0)
        _ffi.check(self._lib.lapper_set_cov(self._h, C.byref(out)))
        return int(out.value)

    def union_and_intersect(self, other: "Lapper") -> Tuple[int, int]:  # src/lib.rs:512-545
        u, i = C.c_uint32(0), C.c_uint32(0)
        _ffi.check(self._lib.lapper_union_and_intersect(self._h, other._h, C.byref(u), C.byref(i)))
        return int(u.value), int(i.value)

    def intersect(self, other: "Lapper") -> int:  # src/lib.rs:550-553
        return self.union_and_intersect(other)[1]

    def union(self, other: "Lapper") -> int:  # src/lib.rs:556-559
        return self.union_and_intersect(other)[0]


    def depth(self) -> List[Tuple[int, int, int]]:
        """src/lib.rs:576-598 + IterDepth: (start, stop, depth) runs inside the merged intervals, in order."""
        n = C.c_uint64(0)
        ps, pe, pd = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        rc = self._lib.lapper_depth(self._h, C.byref(n), C.byref(ps), C.byref(pe), C.byref(pd))
        if rc == _ffi.ERR_INVALID_ARG and len(self) == 0:
            raise IndexError("depth() on an empty lapper panics in the reference (src/lib.rs:744)")
        _ffi.check(rc)
        try:
            k = int(n.value)
            cols = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(k,)).tolist() if k else [] for p in (ps, pe, pd)]
        finally:
            for p in (ps, pe, pd):
                self._lib.lapper_host_free(p)
        return list(zip(*cols)) if k else []


class LapperGroup:
    """One replica of the index per GPU of this process; queries sharded contiguously
    (device g owns [g*nq/G, (g+1)*nq/G)), results identical to the single-GPU call."""

    def __init__(self, lapper: Lapper, devices: Sequence[int]):
        self._lib = _ffi.lib()
        self._h = C.c_void_p(0)
        devs = (C.c_int * len(devices))(*devices)
        _ffi.check(self._lib.lapper_replicate(lapper.handle, devs, len(devices), C.byref(self._h)))

    def __len__(self) -> int:
        return int(self._lib.lapper_group_size(self._h))

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), C.c_void_p(0)
        if h:
            self._lib.lapper_group_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count_batch(self, q_start, q_stop) -> np.ndarray:
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        out = np.empty(qs.size, np.uint64)
        _ffi.check(self._lib.lapper_group_count_batch_u32(self._h, _ptr(qs), _ptr(qe), qs.size, _ptr(out)))
        return out

    def find_batch(self, q_start, q_stop) -> Tuple[np.ndarray, np.ndarray]:
        qs, qe = _u32(q_start), _u32(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        offsets = np.empty(qs.size + 1, np.uint64)
        idx_p, total = C.c_void_p(0), C.c_uint64(0)
        _ffi.check(self._lib.lapper_group_find_batch_u32(self._h, _ptr(qs), _ptr(qe), qs.size, _ptr(offsets),
                                                         C.byref(idx_p), C.byref(total)))
        try:
            n = int(total.value)
            indices = np.ctypeslib.as_array(C.cast(idx_p, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        finally:
            self._lib.lapper_host_free(idx_p)
        return offsets, indices

    # ---------------------------------------------------------------- set operations over the group
    def cov(self) -> int:
        """cov as the sum of the replicas' chunk sums (replica g covers positions [g n/G, (g+1) n/G) of the sorted vector)"""
        out = C.c_uint32(0)
        _ffi.check(self._lib.lapper_group_cov(self._h, C.byref(out)))
        return int(out.value)

    def set_cov(self) -> int:
        out = C.c_uint32(0)
        _ffi.check(self._lib.lapper_group_set_cov(self._h, C.byref(out)))
        return int(out.value)

    def merge_overlaps(self) -> None:
        """every replica merges its own copy, all at once; the replicas stay identical"""
        _ffi.check(self._lib.lapper_group_merge_overlaps(self._h))

    def replica_intervals(self, g: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(starts, stops, perm) of replica g's sorted vector (a borrowed handle: the group owns it)"""
        h = self._lib.lapper_group_replica(self._h, g)
        if not h:
            raise IndexError(g)
        n = int(self._lib.lapper_len(h))
        s, e, p = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint32)
        _ffi.check(self._lib.lapper_get_intervals(h, _ptr(s), _ptr(e), _ptr(p)))
        return s, e, p
