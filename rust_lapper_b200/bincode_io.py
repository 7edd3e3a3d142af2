"""On-disk compatibility with the reference's `with_serde` feature (src/lib.rs:85-86, :90, :104; test :1387-1409):
`bincode::serialize(&lapper)` / `bincode::deserialize` of a `Lapper<I, T>`.

bincode 1.3 with default options (Cargo.lock pins bincode 1.3.3) is little-endian, fixed-width integers, `Vec<X>` = u64
length + elements, `usize` = u64, `Option<X>` = one tag byte (0 = None, 1 = Some) + X, `bool` = one byte, structs =
their fields in declaration order.  The fields of `Lapper` (src/lib.rs:111-122):

    intervals: Vec<Interval<I, T>>   each (start: I, stop: I, val: T)      src/lib.rs:97-99
    starts: Vec<I>, stops: Vec<I>, max_len: I, cov: Option<I>, overlaps_merged: bool

`coord` names I ("u32", "u64"/"usize", "u16", "u8"), `val` names T (an integer type or "bool").  The codec works on a
plain dict, so it is testable without a GPU; `dumps` / `loads` connect it to a `Lapper` / `Lapper64` handle
(include/lapper_b200.hpp carries the same codec for the C++ facade: `to_bincode` / `from_bincode`).  There is no Rust
toolchain in this image, so the byte layout is pinned by the format description above, not by bytes the crate wrote.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Any, Dict

import numpy as np

from . import _ffi

_INT = {"u8": ("<B", 1), "u16": ("<H", 2), "u32": ("<I", 4), "u64": ("<Q", 8), "usize": ("<Q", 8),
        "i8": ("<b", 1), "i16": ("<h", 2), "i32": ("<i", 4), "i64": ("<q", 8), "isize": ("<q", 8), "bool": ("<?", 1)}
_NP = {"u8": np.uint8, "u16": np.uint16, "u32": np.uint32, "u64": np.uint64, "usize": np.uint64}


def encode(state: Dict[str, Any], coord: str = "u32", val: str = "u32") -> bytes:
    """state: intervals = [(start, stop, val)], starts, stops, max_len, cov (None or int), overlaps_merged"""
    cf, vf = _INT[coord][0], _INT[val][0]
    out = [struct.pack("<Q", len(state["intervals"]))]
    for s, e, v in state["intervals"]:
        out.append(struct.pack(cf, s) + struct.pack(cf, e) + struct.pack(vf, v))
    for key in ("starts", "stops"):
        arr = np.asarray(state[key], dtype=_NP[coord])
        out.append(struct.pack("<Q", arr.size) + arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    out.append(struct.pack(cf, state["max_len"]))
    out.append(b"\x00" if state["cov"] is None else b"\x01" + struct.pack(cf, state["cov"]))
    out.append(b"\x01" if state["overlaps_merged"] else b"\x00")
    return b"".join(out)


def decode(buf: bytes, coord: str = "u32", val: str = "u32") -> Dict[str, Any]:
    (cf, cw), (vf, vw) = _INT[coord], _INT[val]
    pos = 0

    def take(fmt, width):
        nonlocal pos
        if pos + width > len(buf):
            raise ValueError("bincode: unexpected end of input")
        (x,) = struct.unpack_from(fmt, buf, pos)
        pos += width
        return x

    n = take("<Q", 8)
    rec = np.dtype([("s", _NP[coord]), ("e", _NP[coord]), ("v", np.dtype(vf.replace("?", "u1")))]).newbyteorder("<")
    if rec.itemsize != 2 * cw + vw or pos + n * rec.itemsize > len(buf):
        raise ValueError("bincode: unexpected end of input")
    ivs = np.frombuffer(buf, dtype=rec, count=n, offset=pos)
    pos += n * rec.itemsize
    vals = [bool(v) for v in ivs["v"]] if val == "bool" else ivs["v"].tolist()
    state: Dict[str, Any] = {"intervals": list(zip(ivs["s"].tolist(), ivs["e"].tolist(), vals))}
    for key in ("starts", "stops"):
        m = take("<Q", 8)
        if pos + m * cw > len(buf):
            raise ValueError("bincode: unexpected end of input")
        state[key] = np.frombuffer(buf, dtype=np.dtype(_NP[coord]).newbyteorder("<"), count=m, offset=pos).astype(_NP[coord])
        pos += m * cw
    state["max_len"] = take(cf, cw)
    tag = take("<B", 1)
    if tag not in (0, 1):
        raise ValueError("bincode: invalid Option tag")
    state["cov"] = take(cf, cw) if tag else None
    flag = take("<B", 1)
    if flag not in (0, 1):
        raise ValueError("bincode: invalid bool")
    state["overlaps_merged"] = bool(flag)
    if pos != len(buf):
        raise ValueError("bincode: trailing bytes")
    return state


def _is_wide(lapper) -> bool:
    from .lapper64 import Lapper64

    return isinstance(lapper, Lapper64)


def dumps(lapper, coord: str = "u32", val: str = "u32", default_val: Any = 0) -> bytes:
    """`bincode::serialize(&lapper)` for a b200 `Lapper` or `Lapper64` (vals come from the host-side list;
    `default_val` without one).  A coordinate that does not fit `coord` is an error, as it could not exist in a
    `Lapper<coord, T>`."""
    lib = _ffi.lib()
    s, e, perm = lapper.intervals_soa()
    vals = lapper._vals
    ivs = [(int(a), int(b), default_val if vals is None else vals[int(p)]) for a, b, p in zip(s.tolist(), e.tolist(), perm.tolist())]
    some, merged = C.c_int(0), C.c_int(0)
    if _is_wide(lapper):
        cov = C.c_uint64(0)
        _ffi.check(lib.lapper64_get_cached_state(lapper._h, C.byref(some), C.byref(cov), C.byref(merged)))
    else:
        cov = C.c_uint32(0)
        _ffi.check(lib.lapper_get_cached_state(lapper.handle, C.byref(some), C.byref(cov), C.byref(merged)))
    try:
        return encode({"intervals": ivs, "starts": lapper.starts, "stops": lapper.stops, "max_len": lapper.max_len,
                       "cov": int(cov.value) if some.value else None, "overlaps_merged": bool(merged.value)}, coord, val)
    except struct.error as err:
        raise ValueError(f"value does not fit the serialised type ({coord} coordinates, {val} vals): {err}") from None


def loads(buf: bytes, coord: str = "u32", val: str = "u32", device: int = 0, engine: str = "auto"):
    """`bincode::deserialize::<Lapper<I, T>>` into a b200 handle.  `engine`: "u32" -> `Lapper`, "u64" -> `Lapper64`,
    "auto" -> `Lapper` when every coordinate fits 32 bits (any width of `coord`), else `Lapper64`.  The index is
    rebuilt from `intervals` -- it is derived data -- and checked against the serialised `starts` / `stops` /
    `max_len`."""
    from .lapper import Lapper
    from .lapper64 import Lapper64

    if engine not in ("auto", "u32", "u64"):
        raise ValueError("engine must be 'auto', 'u32' or 'u64'")
    st = decode(buf, coord, val)
    ivs = st["intervals"]
    beyond = any(a > 0xFFFFFFFF or b > 0xFFFFFFFF for a, b, _ in ivs)
    if engine == "u32" and beyond:
        raise ValueError("coordinates beyond u32 are not supported by the u32 engine")
    wide = engine == "u64" or (engine == "auto" and beyond)
    cls, dt = (Lapper64, np.uint64) if wide else (Lapper, np.uint32)
    lap = cls(starts=np.array([a for a, _, _ in ivs], dtype=dt), stops=np.array([b for _, b, _ in ivs], dtype=dt),
              vals=[v for _, _, v in ivs], device=device)
    if not (np.array_equal(lap.starts, np.asarray(st["starts"], dtype=dt))
            and np.array_equal(lap.stops, np.asarray(st["stops"], dtype=dt)) and lap.max_len == st["max_len"]):
        raise ValueError("serialised starts / stops / max_len do not belong to the serialised intervals")
    some, cov, merged = 0 if st["cov"] is None else 1, st["cov"] or 0, 1 if st["overlaps_merged"] else 0
    if wide:
        _ffi.check(_ffi.lib().lapper64_restore_cached_state(lap._h, some, cov, merged))
    else:
        _ffi.check(_ffi.lib().lapper_restore_cached_state(lap.handle, some, cov, merged))
    return lap
