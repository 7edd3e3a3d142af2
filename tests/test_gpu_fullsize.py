"""BASELINE.json's full sizes on the GPU (device-resident, through the `_dev` C-ABI entry points), checked
through size-independent properties -- the oracle cannot enumerate 100 M queries in test time:

  * count: sum(count) == number of CSR hits find produces for the same queries (the reference's own
    cross-check idiom, find(..).count() == count(..), src/lib.rs:1028 ff.);
  * find: offsets start at 0, are non-decreasing, end at the hit total; every query's slice is strictly
    ascending (iterator order); every emitted position satisfies the overlap predicate (src/lib.rs:140-142);
  * a random sample of queries is compared bit for bit with the oracle;
  * cfg 4 (pathological long intervals): count is unaffected by max_len -- same checks on a sample.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_IV = 10_000_000
NQ = 100_000_000


def _vp(t):
    return C.c_void_p(t.data_ptr())


def _build(lib, s, e, dev):
    import torch

    from rust_lapper_b200 import _ffi

    h = C.c_void_p(0)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _ffi.check(lib.lapper_build_u32_dev(_vp(s), _vp(e), s.numel(), dev.index, 0, sp, C.byref(h)))
    return h, sp


def _count(lib, h, qs, qe, sp):
    import torch

    from rust_lapper_b200 import _ffi

    out = torch.empty(qs.numel(), dtype=torch.int32, device=qs.device)
    _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(qs), _vp(qe), qs.numel(), _vp(out), sp))
    torch.cuda.synchronize()
    return out


def _find(lib, h, qs, qe, sp):
    import torch

    from rust_lapper_b200 import _ffi

    nq = qs.numel()
    off = torch.empty(nq + 1, dtype=torch.int64, device=qs.device)
    total = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(off), C.byref(total), sp))
    idx = torch.empty(max(int(total.value), 1), dtype=torch.int32, device=qs.device)
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(off), _vp(idx), sp))
    torch.cuda.synchronize()
    return off, idx[: int(total.value)], int(total.value)


def _check_csr_properties(torch, view, qs, qe, off, idx, total):
    nq = qs.numel()
    assert int(off[0]) == 0 and int(off[-1]) == total
    assert bool((off[1:] >= off[:-1]).all())
    # strictly ascending inside every query slice: a descent may only happen at a slice boundary
    if total > 1:
        descents = torch.nonzero(idx[1:] <= idx[:-1]).flatten() + 1  # positions p with idx[p] <= idx[p-1]
        if descents.numel():
            is_boundary = torch.zeros(total + 1, dtype=torch.bool, device=idx.device)
            is_boundary[off] = True
            assert bool(is_boundary[descents].all()), "positions not ascending inside a query's slice"
    # every emitted position overlaps its query: start_i < q_stop and stop_i > q_start
    owner = torch.repeat_interleave(torch.arange(nq, device=idx.device), (off[1:] - off[:-1]))
    to_i64 = lambda t: t.to(torch.int64) & 0xFFFFFFFF  # noqa: E731  (u32 stored as int32)
    s = to_i64(view["starts"][idx.to(torch.int64)])
    e = to_i64(view["stops"][idx.to(torch.int64)])
    assert bool(((s < to_i64(qe)[owner]) & (e > to_i64(qs)[owner])).all())


@pytest.mark.parametrize("cfg", ["cfg2_random", "cfg3_sorted"])
def test_full_size_properties(oracle, cfg):
    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    U = synth.GRCH38
    if cfg == "cfg2_random":
        s, e = synth.intervals_uniform(42, N_IV, U, 50, 5000, device=dev)
        qs, qe = synth.queries_fixed_len(43, NQ, U, 150, device=dev)
        hs, he = synth.intervals_uniform(42, N_IV, U, 50, 5000)
    else:
        s, e = synth.intervals_gene_exon(44, N_IV, U, device=dev)
        qs, qe = synth.queries_sorted_fixed_len(45, NQ, U, 150, device=dev)
        hs, he = synth.intervals_gene_exon(44, N_IV, U)
    h, sp = _build(lib, s, e, dev)
    try:
        counts = _count(lib, h, qs, qe, sp)
        off, idx, total = _find(lib, h, qs, qe, sp)
        assert int(counts.to(torch.int64).sum()) == total  # find(..).count() == count(..), summed
        assert bool(((off[1:] - off[:-1]) == (counts.to(torch.int64) & 0xFFFFFFFF)).all())
        # the engine's own sorted arrays, straight from the handle (device view)
        v = _ffi.DevView()
        _ffi.check(lib.lapper_get_dev_view(h, C.byref(v)))
        n = int(v.n)
        assert n == N_IV
        host_s, host_e = np.empty(n, np.uint32), np.empty(n, np.uint32)
        _ffi.check(lib.lapper_get_intervals(h, C.c_void_p(host_s.ctypes.data), C.c_void_p(host_e.ctypes.data), None))
        assert np.all(np.diff(host_s.astype(np.int64)) >= 0)  # sortedness of the build
        view = {"starts": torch.from_numpy(host_s.view(np.int32)).to(dev), "stops": torch.from_numpy(host_e.view(np.int32)).to(dev)}
        _check_csr_properties(torch, view, qs, qe, off, idx, total)
        # a sample against the oracle, bit for bit
        rng = np.random.default_rng(1)
        pick = np.sort(rng.choice(NQ, 200_000, replace=False))
        pick_t = torch.from_numpy(pick).to(dev)
        sq, se = synth.to_numpy_u32(qs[pick_t]), synth.to_numpy_u32(qe[pick_t])
        cpu = oracle.OracleLapper(starts=hs, stops=he)
        assert np.array_equal(cpu.intervals_soa()[0], host_s) and np.array_equal(cpu.intervals_soa()[1], host_e)
        want_counts = cpu.count_batch(sq, se, 8)
        assert np.array_equal(counts[pick_t].cpu().numpy().view(np.uint32).astype(np.uint64), want_counts)
        woff, widx = cpu.find_batch(sq, se, 8)
        off_c = off.cpu().numpy()
        idx_c = idx.cpu().numpy().view(np.uint32)
        for k in rng.choice(pick.size, 2000, replace=False):
            j = int(pick[k])
            assert np.array_equal(idx_c[off_c[j]:off_c[j + 1]], widx[int(woff[k]):int(woff[k + 1])])
    finally:
        lib.lapper_free(h)


def test_pathological_count_and_find_sample(oracle):
    """cfg 4: 1 % of intervals span > 50 % of the universe (max_len ~ 0.9 U).  count at 100 M queries is
    unaffected (BITS); find is checked on a sample (each query has ~70 k hits, so the batch is small)."""
    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    U = synth.GRCH38
    n = 2_000_000
    s, e = synth.intervals_pathological(46, n, U, device=dev)
    hs, he = synth.intervals_pathological(46, n, U)
    qs, qe = synth.queries_fixed_len(47, NQ, U, 150, device=dev)
    h, sp = _build(lib, s, e, dev)
    try:
        counts = _count(lib, h, qs, qe, sp)
        cpu = oracle.OracleLapper(starts=hs, stops=he)
        assert cpu.max_len > U // 2
        m = 300_000
        want = cpu.count_batch(synth.to_numpy_u32(qs[:m]), synth.to_numpy_u32(qe[:m]), 8)
        assert np.array_equal(counts[:m].cpu().numpy().view(np.uint32).astype(np.uint64), want)
        k = 2_000
        off, idx, total = _find(lib, h, qs[:k].contiguous(), qe[:k].contiguous(), sp)
        assert total == int(want[:k].sum())
        woff, widx = cpu.find_batch(synth.to_numpy_u32(qs[:k]), synth.to_numpy_u32(qe[:k]), 8)
        assert np.array_equal(off.cpu().numpy().astype(np.uint64), woff)
        assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx)
    finally:
        lib.lapper_free(h)


def test_tile_with_more_than_2_pow_32_hits():
    """One 1024-query tile whose hits exceed 2^32 (4.4 M identical intervals x 1000 queries inside them), followed in the
    same tile by light queries on a distant cluster: the tile emit kernel's 32-bit tile-relative offsets cannot
    address that, so the tile must go to the 64-bit path; every slice is checked on the device."""
    import torch

    from rust_lapper_b200 import _ffi

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    n_big, n_small = 4_400_000, 64
    s = torch.cat([torch.full((n_big,), 1000, dtype=torch.int32, device=dev),
                   torch.arange(n_small, dtype=torch.int32, device=dev) * 10 + 1_000_000])
    e = torch.cat([torch.full((n_big,), 2000, dtype=torch.int32, device=dev),
                   torch.arange(n_small, dtype=torch.int32, device=dev) * 10 + 1_000_040])
    nq_big, nq = 1000, 3000
    qs = torch.cat([torch.full((nq_big,), 1500, dtype=torch.int32, device=dev),
                    (torch.arange(nq - nq_big, dtype=torch.int32, device=dev) % 600) + 1_000_000])
    qe = qs + 7
    h, sp = _build(lib, s, e, dev)
    try:
        off, idx, total = _find(lib, h, qs, qe, sp)
        counts = _count(lib, h, qs, qe, sp)
        assert total > 2**32
        hits = off[1:] - off[:-1]
        assert bool((hits == (counts.to(torch.int64) & 0xFFFFFFFF)).all())
        assert bool((hits[:nq_big] == n_big).all())
        # heavy slices: positions 0 .. n_big-1 in order (the identical intervals sort first)
        big = idx[: nq_big * n_big].view(nq_big, n_big)
        assert bool((big == torch.arange(n_big, dtype=torch.int32, device=dev)).all())
        del big
        # light slices: ascending positions inside the small cluster, each overlapping its query
        lo = int(off[nq_big])
        light = idx[lo:].to(torch.int64)
        owner = torch.repeat_interleave(torch.arange(nq_big, nq, device=dev), hits[nq_big:])
        assert bool((light >= n_big).all())
        ls = (light - n_big) * 10 + 1_000_000
        assert bool(((ls < qe[owner].to(torch.int64)) & (ls + 40 > qs[owner].to(torch.int64))).all())
        # and the counts are what the geometry says: starts 10 apart, 40 long, query 7 long
        q = qs[nq_big:].to(torch.int64) - 1_000_000
        k = torch.arange(n_small, device=dev, dtype=torch.int64)[None, :] * 10
        want = ((k < (q[:, None] + 7)) & (k + 40 > q[:, None])).sum(1)
        assert bool((want == hits[nq_big:]).all())
    finally:
        lib.lapper_free(h)


def test_cfg5_sized_set_auto_selected_tables(oracle):
    """cfg 5's interval set at its real size (50 M): only n >~ 20 M reaches the third branch of choose_dense_shape
    (build.cu) -- bucket width ~200, hundreds of thousands of overflowed sectors -- and `local_span = 0` (query.cu), which
    sends start-sorted batches through the packed path too.  Random, start-sorted, long and inverted queries; a 200 k
    sample against the oracle for count and find, bit for bit; sum(count) == |CSR| on the whole batches."""
    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    U = synth.GRCH38
    n = 50_000_000
    s, e = synth.intervals_uniform(42, n, U, 50, 5000, device=dev)
    hs, he = synth.intervals_uniform(42, n, U, 50, 5000)
    h, sp = _build(lib, s, e, dev)
    del s, e
    try:
        w, ns, no = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        _ffi.check(lib.lapper_packed_table_info(h, C.byref(w), C.byref(ns), C.byref(no)))
        assert w.value and w.value < 400, "the auto-selected packed table of the dense regime was not built"
        assert no.value > 10_000  # the regime the test is about: many overflowed sectors
        cpu = oracle.OracleLapper(starts=hs, stops=he)
        host_s, host_e, host_p = (np.empty(n, np.uint32) for _ in range(3))
        _ffi.check(lib.lapper_get_intervals(h, C.c_void_p(host_s.ctypes.data), C.c_void_p(host_e.ctypes.data),
                                            C.c_void_p(host_p.ctypes.data)))
        cs, ce, cp = cpu.intervals_soa()
        assert np.array_equal(cs, host_s) and np.array_equal(ce, host_e) and np.array_equal(cp, host_p)
        assert int(lib.lapper_max_len(h)) == cpu.max_len
        del host_s, host_e, host_p, cs, ce, cp
        rng = np.random.default_rng(5)
        nq = 20_000_000
        batches = {
            "random": synth.queries_fixed_len(43, nq, U, 150, device=dev),
            "sorted": synth.queries_sorted_fixed_len(45, nq, U, 150, device=dev),
        }
        # long (beyond any margin), zero-length and inverted queries mixed into a random batch
        ms, me = synth.queries_fixed_len(48, nq, U, 150, device=dev)
        kind = torch.arange(nq, device=dev) % 8
        me = torch.where(kind == 1, ms + 3000, me)
        me = torch.where(kind == 2, ms, me)
        me = torch.where(kind == 3, ms - 100, me)
        batches["mixed"] = (ms.contiguous(), me.to(torch.int32).contiguous())
        for name, (qs, qe) in batches.items():
            counts = _count(lib, h, qs, qe, sp)
            pick = np.sort(rng.choice(nq, 200_000, replace=False))
            pick_t = torch.from_numpy(pick).to(dev)
            sq, se = synth.to_numpy_u32(qs[pick_t]), synth.to_numpy_u32(qe[pick_t])
            want = cpu.count_batch(sq, se, 8)
            got = counts[pick_t].cpu().numpy().view(np.uint32).astype(np.uint64)
            assert np.array_equal(got, want & 0xFFFFFFFF), name
            # find on a prefix (43 hits per query: 2 M queries ~ 86 M positions)
            k = 2_000_000
            off, idx, total = _find(lib, h, qs[:k].contiguous(), qe[:k].contiguous(), sp)
            if name != "mixed":  # |find| == count needs non-empty queries
                assert total == int((counts[:k].to(torch.int64) & 0xFFFFFFFF).sum()), name
            sub = np.sort(rng.choice(k, 100_000, replace=False))
            fq, fe = synth.to_numpy_u32(qs[:k])[sub], synth.to_numpy_u32(qe[:k])[sub]
            woff, widx = cpu.find_batch(fq, fe, 8)
            off_c = off.cpu().numpy()
            idx_c = idx.cpu().numpy().view(np.uint32)
            hits = (off_c[1:] - off_c[:-1])[sub]
            assert np.array_equal(hits.astype(np.uint64), np.diff(woff)), name
            for t in rng.choice(sub.size, 3000, replace=False):
                j = int(sub[t])
                assert np.array_equal(idx_c[off_c[j]:off_c[j + 1]], widx[int(woff[t]):int(woff[t + 1])]), name
            del counts, off, idx
        # merge_overlaps + cov at this size, and a count against the merged set (lazy tables)
        cov = C.c_uint32(0)
        _ffi.check(lib.lapper_cov(h, C.byref(cov)))
        assert int(cov.value) == cpu.cov()
        _ffi.check(lib.lapper_merge_overlaps(h))
        cpu.merge_overlaps()
        m = int(lib.lapper_len(h))
        assert m == len(cpu)
        ms_, me_ = np.empty(m, np.uint32), np.empty(m, np.uint32)
        _ffi.check(lib.lapper_get_intervals(h, C.c_void_p(ms_.ctypes.data), C.c_void_p(me_.ctypes.data), None))
        assert np.array_equal(ms_, cpu.intervals_soa()[0]) and np.array_equal(me_, cpu.intervals_soa()[1])
        assert int(lib.lapper_max_len(h)) == cpu.max_len
        qs, qe = batches["random"]
        counts = _count(lib, h, qs[:1_000_000].contiguous(), qe[:1_000_000].contiguous(), sp)
        want = cpu.count_batch(synth.to_numpy_u32(qs[:1_000_000]), synth.to_numpy_u32(qe[:1_000_000]), 8)
        assert np.array_equal(counts.cpu().numpy().view(np.uint32).astype(np.uint64), want)
    finally:
        lib.lapper_free(h)


def test_north_star_10m_intervals_x_1b_queries(oracle):
    """BASELINE.json's target size: 10 M intervals x 1 B queries, bit-exact.  The oracle cannot count a billion queries
    in test time, so the billion is covered by two independent engines agreeing: ten batches of 100 M random-order
    queries are counted by the packed-table kernel, then sorted on the device and counted again by the sorted-run
    kernel (a different algorithm over different arrays: ranks against the sorted starts / stops instead of one table
    sector per query) -- the two answers must be equal query for query -- and find's CSR over the sorted batch must
    have exactly those counts as its row lengths (one-call device API).  20 k queries of every batch are compared with
    the oracle itself, so 200 k oracle answers anchor the agreement."""
    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    U = synth.GRCH38
    s, e = synth.intervals_uniform(42, N_IV, U, 50, 5000, device=dev)
    hs, he = synth.intervals_uniform(42, N_IV, U, 50, 5000)
    cpu = oracle.OracleLapper(starts=hs, stops=he)
    h, sp = _build(lib, s, e, dev)
    stats = (C.c_uint64 * 2)()
    try:
        lib.lapper_debug_sorted_run_stats(stats, 1)
        grand = 0
        for batch in range(10):
            qs, qe = synth.queries_fixed_len(1000 + batch, NQ, U, 150, device=dev)
            c_random = _count(lib, h, qs, qe, sp)
            order = torch.argsort(qs.to(torch.int64) & 0xFFFFFFFF)
            qs_s, qe_s = qs[order].contiguous(), qe[order].contiguous()
            c_sorted = _count(lib, h, qs_s, qe_s, sp)
            assert bool(torch.equal(c_sorted, c_random[order])), f"batch {batch}: the two count engines disagree"
            # find over the sorted batch: row lengths == counts, total == their sum
            total = int(c_sorted.to(torch.int64).sum())
            off = torch.empty(NQ + 1, dtype=torch.int64, device=dev)
            idx = torch.empty(total, dtype=torch.int32, device=dev)
            t = C.c_uint64(0)
            _ffi.check(lib.lapper_find_batch_u32_dev(h, _vp(qs_s), _vp(qe_s), NQ, _vp(off), _vp(idx), total, C.byref(t), sp))
            torch.cuda.synchronize()
            assert int(t.value) == total
            assert bool(torch.equal(off[1:] - off[:-1], c_sorted.to(torch.int64)))
            grand += total
            # the oracle on a sample of this batch (random-order positions)
            pick = torch.randint(0, NQ, (20_000,), device=dev, generator=torch.Generator(device=dev).manual_seed(batch))
            a = qs[pick].cpu().numpy().view(np.uint32)
            b = qe[pick].cpu().numpy().view(np.uint32)
            want = cpu.count_batch(a, b, 4)
            assert np.array_equal(c_random[pick].cpu().numpy().view(np.uint32).astype(np.uint64), want)
            # ... and find's rows for the same queries in the sorted batch
            inv = torch.empty_like(order)
            inv[order] = torch.arange(NQ, device=dev)
            rows = inv[pick]
            wo, wi = cpu.find_batch(a, b, 4)
            lo = off[rows].cpu().numpy()
            got = np.concatenate([idx[int(x): int(x) + int(n)].cpu().numpy().view(np.uint32) for x, n in
                                  zip(lo[:2000], np.diff(wo)[:2000])]) if len(lo) else np.zeros(0, np.uint32)
            assert np.array_equal(got, wi[: int(wo[2000])])
            del qs, qe, qs_s, qe_s, c_random, c_sorted, off, idx, order, inv
        lib.lapper_debug_sorted_run_stats(stats, 0)
        assert stats[0] >= 10 * (NQ // 256) * 2 - 10 * 8  # count + find's count pass went through the sorted-run kernel
        assert grand > 0
    finally:
        lib.lapper_free(h)
