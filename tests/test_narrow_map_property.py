"""The coordinate map of the u64 engine's narrowing (csrc/engine64.cu, `NarrowMap`), restated and checked exhaustively
on a small word size: with W-bit "u32" and 2W-bit "u64", for every set range [lo, hi] the map accepts and every query
coordinate q, the three comparisons the reference makes against a set coordinate x (src/lib.rs:140-142, :611-618:
x < q, x <= q, x > q) have the same outcome on the narrowed values.  This pins the argument in DESIGN.md §4; the kernels
themselves are pinned by tests/test_gpu_u64.py against the oracle."""
import itertools

W = 4                       # the narrow type has W bits, the wide one 2W
NARROW_MAX = (1 << W) - 1   # u32::MAX
WIDE_MAX = (1 << (2 * W)) - 1
MAX_COORD = NARROW_MAX - 2  # kNarrowMaxCoord = 2^32 - 3
MAX_SPAN = NARROW_MAX - 3   # kNarrowMaxSpan  = 2^32 - 4


def narrow_map_of(lo, hi):
    """(base, add, top) as narrow_map_of() chooses them, or None"""
    if hi <= MAX_COORD:
        return 0, 0, NARROW_MAX - 1
    if hi - lo <= MAX_SPAN:
        return lo, 1, hi - lo + 2
    return None


def m(q, base, add, top):
    if q < base:
        return 0
    d = q - base
    return top if d >= top - add else d + add


def test_every_comparison_survives_the_map():
    checked = 0
    for lo, hi in itertools.combinations_with_replacement(range(WIDE_MAX + 1), 2):
        mp = narrow_map_of(lo, hi)
        if mp is None:
            continue
        xs = sorted({lo, hi, (lo + hi) // 2, min(lo + 1, hi), max(hi - 1, lo)})
        fx = [m(x, *mp) for x in xs]
        assert all(0 <= v <= NARROW_MAX - 2 for v in fx), (lo, hi)           # set coordinates stay <= 2^32 - 3
        assert all(a < b for a, b in zip(fx, fx[1:])), (lo, hi)              # and keep their order
        for q in range(WIDE_MAX + 1):
            fq = m(q, *mp)
            assert fq <= NARROW_MAX - 1                                       # u32::MAX stays free for the start + 1 wrap
            for x, v in zip(xs, fx):
                assert (x < q) == (v < fq) and (x <= q) == (v <= fq) and (x > q) == (v > fq), (lo, hi, x, q)
            checked += 1
    assert checked > 10_000


def test_sets_that_do_not_fit_are_refused():
    assert narrow_map_of(0, MAX_COORD) == (0, 0, NARROW_MAX - 1)
    assert narrow_map_of(0, MAX_COORD + 1) is None                            # starts at 0, one coordinate too high
    assert narrow_map_of(5, 5 + MAX_SPAN) == (5, 1, MAX_SPAN + 2)
    assert narrow_map_of(5, 5 + MAX_SPAN + 1) is None
    assert narrow_map_of(WIDE_MAX - MAX_SPAN, WIDE_MAX) == (WIDE_MAX - MAX_SPAN, 1, MAX_SPAN + 2)
