"""`Lapper<u64, T>`: the host-side mirror of the reference's Lapper (src/lib.rs:182-680) over the u64 entry points
of include/lapper_b200.h (csrc/engine64.cu).  Same method names and meaning as `rust_lapper_b200.Lapper`; `val`
stays on the host, results are positions in the sorted interval vector."""
from __future__ import annotations

import ctypes as C
from typing import Any, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi

U64_MAX = (1 << 64) - 1


def _u64(a) -> np.ndarray:
    """Coordinates as a contiguous u64 array; values outside [0, 2^64) are an error, never wrapped."""
    if isinstance(a, np.ndarray) and a.dtype == np.uint64:
        return np.ascontiguousarray(a).reshape(-1)
    if isinstance(a, np.ndarray) and a.dtype.kind in "iub":
        if a.size and a.dtype.kind == "i" and a.min() < 0:
            raise ValueError("coordinate out of range for u64 (negative)")
        return np.ascontiguousarray(a.astype(np.uint64)).reshape(-1)
    vals = [int(x) for x in np.asarray(a, dtype=object).reshape(-1)]
    if any(x < 0 or x > U64_MAX for x in vals):
        raise ValueError("coordinate out of range for u64 (0 <= x <= 2^64 - 1)")
    return np.array(vals, dtype=np.uint64)


def _ptr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class Lapper64:
    """`Lapper<u64, T>` on one B200.  `intervals` are (start, stop[, val]) tuples or, with `starts=`/`stops=`, two
    integer arrays (vals optional)."""

    def __init__(self, intervals: Iterable[Sequence] = (), *, starts=None, stops=None, vals=None, device: int = 0):
        self._lib = _ffi.lib()
        self._h = C.c_void_p(0)
        if starts is None:
            ivs = list(intervals)
            starts = [iv[0] for iv in ivs]
            stops = [iv[1] for iv in ivs]
            if ivs and len(ivs[0]) > 2:
                vals = [iv[2] for iv in ivs]
        s, e = _u64(starts), _u64(stops)
        if s.shape != e.shape:
            raise ValueError("starts and stops differ in length")
        self._vals = list(vals) if vals is not None else None
        _ffi.check(self._lib.lapper64_build(_ptr(s), _ptr(e), s.size, device, C.byref(self._h)))

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), C.c_void_p(0)
        if h:
            self._lib.lapper64_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- structure
    def __len__(self) -> int:  # src/lib.rs:284-287
        return int(self._lib.lapper64_len(self._h))

    def len(self) -> int:
        return len(self)

    def is_empty(self) -> bool:  # src/lib.rs:296-299
        return len(self) == 0

    @property
    def max_len(self) -> int:
        return int(self._lib.lapper64_max_len(self._h))

    @property
    def overlaps_merged(self) -> bool:
        return bool(self._lib.lapper64_overlaps_merged(self._h))

    def tune_query_length(self, typical_query_len: int) -> None:
        """see Lapper.tune_query_length; takes effect where the u32 engine answers (coordinates that fit 32 bits)"""
        _ffi.check(self._lib.lapper64_tune_query_length(self._h, int(typical_query_len)))

    @property
    def is_narrow(self) -> bool:
        """True once this handle's batches are answered by the u32 engine (every coordinate fits 32 bits)."""
        return bool(_ffi.lib().lapper64_is_narrow(self._h))

    @property
    def has_inverted(self) -> bool:
        return bool(self._lib.lapper64_has_inverted(self._h))

    def intervals_soa(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(starts, stops, input positions) of the sorted `intervals` vector"""
        n = len(self)
        s, e, p = np.empty(n, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint32)
        _ffi.check(self._lib.lapper64_get_intervals(self._h, _ptr(s), _ptr(e), _ptr(p)))
        return s, e, p

    @property
    def intervals(self) -> List[Tuple[int, int]]:
        s, e, _ = self.intervals_soa()
        return list(zip(s.tolist(), e.tolist()))

    def iter(self):  # src/lib.rs:353-359
        return iter(self.intervals)

    @property
    def starts(self) -> np.ndarray:  # private `starts` (src/lib.rs:114): the sorted vector's starts
        return self.intervals_soa()[0]

    @property
    def stops(self) -> np.ndarray:  # private sorted `stops` (src/lib.rs:116)
        out = np.empty(len(self), np.uint64)
        _ffi.check(self._lib.lapper64_get_sorted_stops(self._h, _ptr(out)))
        return out

    def val_at(self, pos: int) -> Any:
        _, _, p = self.intervals_soa()
        ip = int(p[pos])
        return self._vals[ip] if self._vals is not None else ip

    # ---------------------------------------------------------------- batch queries
    def count_batch(self, q_start, q_stop) -> np.ndarray:
        qs, qe = _u64(q_start), _u64(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        out = np.empty(qs.size, np.uint64)
        _ffi.check(self._lib.lapper64_count_batch(self._h, _ptr(qs), _ptr(qe), qs.size, _ptr(out)))
        return out

    def find_batch(self, q_start, q_stop) -> Tuple[np.ndarray, np.ndarray]:
        """CSR: indices[offsets[j]:offsets[j+1]] are the positions find(q_j) yields, in order."""
        qs, qe = _u64(q_start), _u64(q_stop)
        if qs.shape != qe.shape:
            raise ValueError("q_start and q_stop differ in length")
        res = C.c_void_p(0)
        _ffi.check(self._lib.lapper64_find_batch(self._h, _ptr(qs), _ptr(qe), qs.size, C.byref(res)))
        try:
            total = int(self._lib.lapper_result_total(res))
            offsets = np.empty(qs.size + 1, np.uint64)
            indices = np.empty(total, np.uint32)
            _ffi.check(self._lib.lapper_result_copy(res, _ptr(offsets), _ptr(indices)))
        finally:
            self._lib.lapper_result_free(res)
        return offsets, indices

    # ---------------------------------------------------------------- per-query API of the reference
    def count(self, start: int, stop: int) -> int:  # src/lib.rs:610-618
        return int(self.count_batch([start], [stop])[0])

    def find_idx(self, start: int, stop: int) -> List[int]:
        return self.find_batch([start], [stop])[1].tolist()

    def find(self, start: int, stop: int) -> List[Tuple[int, int]]:  # src/lib.rs:628-639
        ivs = self.intervals
        return [ivs[i] for i in self.find_idx(start, stop)]

    def seek_idx(self, start: int, stop: int, cursor: List[int]) -> List[int]:
        """src/lib.rs:656-679: the cursor (a one-element list standing in for `&mut usize`) is updated exactly as the
        reference updates it; the positions are find's from the cursor on."""
        c = C.c_uint64(cursor[0])
        _ffi.check(self._lib.lapper64_seek_cursor(self._h, start, C.byref(c)))
        cursor[0] = int(c.value)
        return [i for i in self.find_idx(start, stop) if i >= cursor[0]]

    def seek(self, start: int, stop: int, cursor: List[int]) -> List[Tuple[int, int]]:
        ivs = self.intervals
        return [ivs[i] for i in self.seek_idx(start, stop, cursor)]

    def insert(self, start: int, stop: int, val: Any = None) -> None:
        """src/lib.rs:260-273: sorted insert (before equal intervals); clears cov and overlaps_merged."""
        pos = C.c_uint64(0)
        _ffi.check(self._lib.lapper64_insert(self._h, start, stop, C.byref(pos)))
        if self._vals is not None:
            while len(self._vals) < int(pos.value):
                self._vals.append(None)
            self._vals.append(val)

    def depth(self) -> List[Tuple[int, int, int]]:
        """src/lib.rs:576-598: maximal runs of equal coverage depth as (start, stop, depth), in iterator order"""
        n = C.c_uint64(0)
        ps, pe, pd = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        _ffi.check(self._lib.lapper64_depth(self._h, C.byref(n), C.byref(ps), C.byref(pe), C.byref(pd)))
        k = int(n.value)
        try:
            s = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_uint64)), shape=(k,)).tolist() if k else []
            e = np.ctypeslib.as_array(C.cast(pe, C.POINTER(C.c_uint64)), shape=(k,)).tolist() if k else []
            d = np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint32)), shape=(k,)).tolist() if k else []
        finally:
            for p in (ps, pe, pd):
                self._lib.lapper_host_free(p)
        return list(zip(s, e, d))

    # ---------------------------------------------------------------- set operations
    def merge_overlaps(self) -> None:  # src/lib.rs:361-416
        _ffi.check(self._lib.lapper64_merge_overlaps(self._h))

    def cov(self) -> int:  # src/lib.rs:310-316
        out = C.c_uint64(0)
        _ffi.check(self._lib.lapper64_cov(self._h, C.byref(out)))
        return int(out.value)

    def set_cov(self) -> int:  # src/lib.rs:320-324
        out = C.c_uint64(0)
        _ffi.check(self._lib.lapper64_set_cov(self._h, C.byref(out)))
        return int(out.value)

    def union_and_intersect(self, other: "Lapper64") -> Tuple[int, int]:  # src/lib.rs:512-545
        u, i = C.c_uint64(0), C.c_uint64(0)
        _ffi.check(self._lib.lapper64_union_and_intersect(self._h, other._h, C.byref(u), C.byref(i)))
        return int(u.value), int(i.value)

    def intersect(self, other: "Lapper64") -> int:  # src/lib.rs:550-553
        return self.union_and_intersect(other)[1]

    def union(self, other: "Lapper64") -> int:  # src/lib.rs:556-559
        return self.union_and_intersect(other)[0]
