"""rust_lapper_b200 -- B200-native batched interval-overlap engine behind rust-lapper's API.

Host-side mirrors of the reference's `Lapper<u32, T>` and `Lapper<u64, T>` (src/lib.rs:182-680) over the C ABI of
include/lapper_b200.h.  All compute runs in hand-written sm_100a CUDA kernels
(rust_lapper_b200/csrc); there is no CPU path.
"""
from ._ffi import LapperError, lib  # noqa: F401
from .lapper import Interval, Lapper, LapperGroup  # noqa: F401
from .lapper64 import Lapper64  # noqa: F401

__all__ = ["Interval", "Lapper", "Lapper64", "LapperGroup", "LapperError", "lib"]
