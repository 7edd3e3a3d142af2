"""The reference's `with_serde` feature (src/lib.rs:85-86; test :1387-1409): bincode 1.3 round trip of a `Lapper`.
The byte layout is checked against a hand-laid-out buffer (format notes in rust_lapper_b200/bincode_io.py: there is no
Rust toolchain here to write real bytes); the reference's own serde_test runs through the engine on the GPU."""
import struct

import numpy as np
import pytest

from rust_lapper_b200 import bincode_io

SERDE_TEST_DATA = [  # src/lib.rs:1389-1397
    (25264912, 25264986, 0), (27273024, 27273065, 0), (27440273, 27440318, 0), (27488033, 27488125, 0),
    (27938410, 27938470, 0), (27959118, 27959171, 0), (28866309, 33141404, 0),
]


def test_layout_of_a_two_interval_lapper_usize_u32():
    # Lapper<usize, u32> { intervals: [(5, 9, 7), (6, 20, 8)], starts: [5, 6], stops: [9, 20], max_len: 14, cov: Some(15),
    #                      overlaps_merged: false }
    want = b"".join([
        struct.pack("<Q", 2),                                   # intervals.len()
        struct.pack("<QQI", 5, 9, 7), struct.pack("<QQI", 6, 20, 8),  # Interval { start, stop, val }
        struct.pack("<Q", 2), struct.pack("<QQ", 5, 6),         # starts
        struct.pack("<Q", 2), struct.pack("<QQ", 9, 20),        # stops
        struct.pack("<Q", 14),                                  # max_len
        b"\x01", struct.pack("<Q", 15),                         # cov: Some(15)
        b"\x00",                                                # overlaps_merged
    ])
    state = {"intervals": [(5, 9, 7), (6, 20, 8)], "starts": [5, 6], "stops": [9, 20], "max_len": 14, "cov": 15,
             "overlaps_merged": False}
    assert bincode_io.encode(state, "usize", "u32") == want
    back = bincode_io.decode(want, "usize", "u32")
    assert back["intervals"] == state["intervals"] and back["max_len"] == 14 and back["cov"] == 15
    assert back["starts"].tolist() == [5, 6] and back["stops"].tolist() == [9, 20] and back["overlaps_merged"] is False


@pytest.mark.parametrize("coord,val", [("u32", "bool"), ("u32", "u32"), ("usize", "u64"), ("u16", "u8")])
def test_codec_round_trip(coord, val):
    rng = np.random.default_rng(3)
    hi = 60000 if coord == "u16" else 10**9
    s = np.sort(rng.integers(0, hi - 100, 50))
    e = s + rng.integers(0, 100, 50)
    vals = [bool(x) for x in rng.integers(0, 2, 50)] if val == "bool" else rng.integers(0, 200, 50).tolist()
    state = {"intervals": list(zip(s.tolist(), e.tolist(), vals)), "starts": s, "stops": np.sort(e),
             "max_len": int((e - s).max()), "cov": None, "overlaps_merged": True}
    raw = bincode_io.encode(state, coord, val)
    back = bincode_io.decode(raw, coord, val)
    assert back["intervals"] == state["intervals"] and back["cov"] is None and back["overlaps_merged"] is True
    assert bincode_io.encode(back, coord, val) == raw


def test_malformed_input_is_rejected():
    state = {"intervals": [(1, 2, 0)], "starts": [1], "stops": [2], "max_len": 1, "cov": None, "overlaps_merged": False}
    raw = bincode_io.encode(state, "u32", "u32")
    for bad in (raw[:-1], raw + b"\x00", raw[:-2] + b"\x02" + raw[-1:], raw[:-1] + b"\x07", b"\xff" * 8):
        with pytest.raises(ValueError):
            bincode_io.decode(bad, "u32", "u32")


@pytest.mark.gpu
def test_reference_serde_test_through_the_engine():
    from rust_lapper_b200 import Lapper

    lapper = Lapper(SERDE_TEST_DATA)
    raw = bincode_io.dumps(lapper, coord="usize", val="u32")           # bincode::serialize(&lapper)
    back = bincode_io.loads(raw, coord="usize", val="u32")             # bincode::deserialize::<Lapper<usize, u32>>
    assert back.find(28974798, 33141355) == [(28866309, 33141404)]     # src/lib.rs:1403-1406
    assert back.count(28974798, 33141355) == 1                         # :1407
    assert back.interval_at(6).val == 0
    # the cached fields survive: cov: Option<I> and overlaps_merged (src/lib.rs:119-122)
    lapper.set_cov()
    lapper.merge_overlaps()
    again = bincode_io.loads(bincode_io.dumps(lapper, "u32", "u32"), "u32", "u32")
    assert again.overlaps_merged and again.cov() == lapper.cov() and again.intervals == lapper.intervals
    state = bincode_io.decode(bincode_io.dumps(lapper, "u32", "u32"), "u32", "u32")
    assert state["cov"] == lapper.cov() and state["overlaps_merged"] is True


@pytest.mark.gpu
def test_serde_through_the_u64_engine():
    """Lapper<usize, u32> with coordinates beyond 32 bits: `loads` picks the u64 engine, cached fields survive."""
    from rust_lapper_b200 import Lapper64

    big = 1 << 40
    data = [(big + a, big + b, v) for a, b, v in SERDE_TEST_DATA]
    lapper = Lapper64(data)
    raw = bincode_io.dumps(lapper, coord="usize", val="u32")
    with pytest.raises(ValueError):
        bincode_io.dumps(lapper, coord="u32", val="u32")               # no Lapper<u32, _> can hold these
    with pytest.raises(ValueError):
        bincode_io.loads(raw, coord="usize", val="u32", engine="u32")
    back = bincode_io.loads(raw, coord="usize", val="u32")
    assert isinstance(back, Lapper64)
    assert back.find(big + 28974798, big + 33141355) == [(big + 28866309, big + 33141404)]
    assert back.count(big + 28974798, big + 33141355) == 1
    lapper.set_cov()
    lapper.merge_overlaps()
    again = bincode_io.loads(bincode_io.dumps(lapper, "u64", "u32"), "u64", "u32")
    assert again.overlaps_merged and again.cov() == lapper.cov() and again.intervals == lapper.intervals
    # a set that fits 32 bits loads into the u32 engine unless the u64 one is asked for
    small = bincode_io.dumps(Lapper64(SERDE_TEST_DATA), coord="usize", val="u32")
    assert not isinstance(bincode_io.loads(small, "usize", "u32"), Lapper64)
    assert isinstance(bincode_io.loads(small, "usize", "u32", engine="u64"), Lapper64)
