// api.cu -- the extern "C" surface declared in include/lapper_b200.h.
// Status codes only; no exception or abort crosses the ABI.  No CPU fallback anywhere: every entry
// point that computes runs the CUDA kernels of build.cu / query.cu / setops.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "index.cuh"

using namespace lap;

namespace lap {
cudaMemPool_t scratch_pool() {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    LAPPER_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) throw Error(LAPPER_ERR_INVALID_ARG, "device index out of range for the scratch pool");
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        LAPPER_CUDA_CHECK(cudaMemPoolCreate(&pools[dev], &props));
        uint64_t keep = ~0ull;
        LAPPER_CUDA_CHECK(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    return pools[dev];
}
}  // namespace lap

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return LAPPER_OK;
    } catch (const Error &e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        g_last_error = "host allocation failed";
        return LAPPER_ERR_OOM;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return LAPPER_ERR_CUDA;
    } catch (...) {
        g_last_error = "unknown error";
        return LAPPER_ERR_CUDA;
    }
}

void require_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw Error(LAPPER_ERR_NO_DEVICE,
                    std::string("no CUDA device available (liblapper_b200 has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= count)
        throw Error(LAPPER_ERR_INVALID_ARG, "device index out of range: " + std::to_string(device));
}

struct StreamOwner {
    cudaStream_t s = nullptr;
    StreamOwner() { LAPPER_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
    ~StreamOwner() {
        if (s) cudaStreamDestroy(s);
    }
};

void check_handle(const lapper_t *l) { LAPPER_REQUIRE(l != nullptr, "null lapper handle"); }

void need_valid_intervals(const lapper_t *l, const char *what) {
    if (l->ix.has_inverted)
        throw Error(LAPPER_ERR_PRECONDITION,
                    std::string(what) + ": the set holds an interval with stop < start; the reference's result is "
                                        "an artefact of arithmetic overflow there and is not reproduced");
}

// chunked host <-> device pipeline for count: three slots, each its own stream, so the H2D of chunk
// i+1, the kernel of chunk i and the D2H of chunk i-1 overlap (PCIe is full duplex).
template <typename OutT>
void count_host(const lapper_t *l, const uint32_t *qs, const uint32_t *qe, uint64_t nq, OutT *out) {
    check_handle(l);
    if (nq == 0) return;
    LAPPER_REQUIRE(qs && qe && out, "count_batch: null pointer");
    DeviceGuard dg(l->device);
    constexpr int kSlots = 3;
    // 8 M queries per chunk: measured best of 2^20 .. 2^25 (20.0 / 18.0 / 17.1 / 16.8 / 17.0 / 18.0 ms per 100 M queries)
    const uint64_t chunk = std::min<uint64_t>(nq, 1ull << 23);
    StreamOwner st[kSlots];
    Scratch<uint32_t> d_qs[kSlots], d_qe[kSlots];
    Scratch<OutT> d_out[kSlots];
    const int used = (int)std::min<uint64_t>(kSlots, div_up(nq, chunk));
    for (int s = 0; s < used; s++) {
        d_qs[s].alloc(chunk, st[s].s);
        d_qe[s].alloc(chunk, st[s].s);
        d_out[s].alloc(chunk, st[s].s);
    }
    uint64_t done = 0;
    for (int i = 0; done < nq; i++) {
        const int s = i % kSlots;
        const uint64_t m = std::min(chunk, nq - done);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qs[s].p, qs + done, m * 4, cudaMemcpyHostToDevice, st[s].s));
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qe[s].p, qe + done, m * 4, cudaMemcpyHostToDevice, st[s].s));
        if (sizeof(OutT) == 8)
            launch_count(l->ix, d_qs[s].p, d_qe[s].p, m, nullptr, reinterpret_cast<uint64_t *>(d_out[s].p), st[s].s);
        else
            launch_count(l->ix, d_qs[s].p, d_qe[s].p, m, reinterpret_cast<uint32_t *>(d_out[s].p), nullptr, st[s].s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(out + done, d_out[s].p, m * sizeof(OutT), cudaMemcpyDeviceToHost, st[s].s));
        done += m;
    }
    for (int s = 0; s < used; s++) {
        d_qs[s].release();
        d_qe[s].release();
        d_out[s].release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st[s].s));
    }
}

lapper_result_t *find_on_device(const lapper_t *l, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                                cudaStream_t stream) {
    std::unique_ptr<lapper_result_t> r(new lapper_result_t());
    r->device = l->device;
    r->nq = nq;
    try {
        LAPPER_CUDA_CHECK(cudaMalloc((void **)&r->d_offsets, (nq + 1) * sizeof(uint64_t)));
        r->total = launch_find_offsets(l->ix, d_qs, d_qe, nq, r->d_offsets, stream);
        LAPPER_CUDA_CHECK(cudaMalloc((void **)&r->d_indices, std::max<uint64_t>(r->total, 1) * sizeof(uint32_t)));
        launch_find_fill(l->ix, d_qs, d_qe, nq, r->d_offsets, r->d_indices, stream);
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    } catch (...) {
        cudaFree(r->d_offsets);
        cudaFree(r->d_indices);
        throw;
    }
    return r.release();
}

__global__ void seek_cursor_kernel(const uint32_t *__restrict__ S, uint32_t n, uint32_t max_len, uint32_t start,
                                   uint64_t cursor_in, uint64_t *__restrict__ cursor_out) {
    // src/lib.rs:658-671 with the advancing while-loop in closed form: it stops at lb - 1
    const uint32_t floor = start >= max_len ? start - max_len : 0u;
    uint64_t c = cursor_in;
    if (c == 0 || (c < n && S[c] > start)) c = lower_bound_u32(S, 0, n, floor);
    if (c + 1 < n && S[c + 1] < floor) c = (uint64_t)lower_bound_u32(S, 0, n, floor) - 1;
    *cursor_out = c;
}

// positions the reference's insert() picks (src/lib.rs:261-263): lower_bound of the interval among the sorted
// intervals by (start, stop), and of its stop among the sorted stops
__global__ void insert_positions_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ Eiv,
                                        const uint32_t *__restrict__ E, uint32_t n, uint32_t start, uint32_t stop,
                                        uint32_t *__restrict__ out) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {  // first interval that is not < (start, stop)
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint32_t s = S[mid], e = Eiv[mid];
        if (s < start || (s == start && e < stop))
            lo = mid + 1;
        else
            hi = mid;
    }
    out[0] = lo;
    out[1] = lower_bound_u32(E, 0, n, stop);
}

// dst[0..pos) = src[0..pos), dst[pos] = value, dst[pos+1..n+1) = src[pos..n)
void insert_into(uint32_t *dst, const uint32_t *src, uint64_t n, uint64_t pos, uint32_t value, cudaStream_t stream) {
    if (pos) LAPPER_CUDA_CHECK(cudaMemcpyAsync(dst, src, pos * 4, cudaMemcpyDeviceToDevice, stream));
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(dst + pos, &value, 4, cudaMemcpyHostToDevice, stream));
    if (n > pos) LAPPER_CUDA_CHECK(cudaMemcpyAsync(dst + pos + 1, src + pos, (n - pos) * 4, cudaMemcpyDeviceToDevice, stream));
}

void replace_index_with_sorted(lapper_t *l, uint32_t *s, uint32_t *e, uint32_t *perm, uint64_t m,
                               cudaStream_t stream) {
    uint32_t *e_sorted = nullptr;
    LAPPER_CUDA_CHECK(cudaMalloc((void **)&e_sorted, std::max<uint64_t>(m, 1) * 4));
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(e_sorted, e, m * 4, cudaMemcpyDeviceToDevice, stream));
    free_index(l->ix);
    l->ix.n = m;
    l->ix.starts = s;
    l->ix.stops_by_iv = e;
    l->ix.stops_sorted = e_sorted;
    l->ix.perm = perm;
    l->ix.bytes = 4 * 4 * std::max<uint64_t>(m, 1);
    finish_index_from_sorted(l->ix, l->build_flags, stream);
}

}  // namespace

struct lapper_group {
    std::vector<lapper_t *> replicas;
};

extern "C" {

int lapper_version(void) { return 100; }

const char *lapper_last_error(void) { return g_last_error.c_str(); }

int lapper_device_count(int *count_out) {
    return guarded([&] {
        LAPPER_REQUIRE(count_out, "null pointer");
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) {
            cudaGetLastError();
            c = 0;
        }
        *count_out = c;
    });
}

int lapper_build_u32_dev(const uint32_t *d_starts, const uint32_t *d_stops, uint64_t n, int device, uint32_t flags,
                         void *stream, lapper_t **out) {
    return guarded([&] {
        LAPPER_REQUIRE(out, "null output handle");
        *out = nullptr;
        LAPPER_REQUIRE(n == 0 || (d_starts && d_stops), "lapper_build: null interval arrays");
        require_device(device);
        DeviceGuard dg(device);
        std::unique_ptr<lapper_t> l(new lapper_t());
        l->device = device;
        l->build_flags = flags;
        build_index_from_unsorted(l->ix, d_starts, d_stops, n, flags, (cudaStream_t)stream);
        l->next_input_id = n;
        LAPPER_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        *out = l.release();
    });
}

int lapper_build_u32_ex(const uint32_t *starts, const uint32_t *stops, uint64_t n, int device, uint32_t flags,
                        lapper_t **out) {
    return guarded([&] {
        LAPPER_REQUIRE(out, "null output handle");
        *out = nullptr;
        LAPPER_REQUIRE(n == 0 || (starts && stops), "lapper_build: null interval arrays");
        LAPPER_REQUIRE(n < 0xFFFFFFF0ull, "lapper_build: n must be < 2^32 - 16 (positions are u32)");
        require_device(device);
        DeviceGuard dg(device);
        StreamOwner st;
        Scratch<uint32_t> d_s(n, st.s), d_e(n, st.s);
        if (n) {
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_s.p, starts, n * 4, cudaMemcpyHostToDevice, st.s));
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_e.p, stops, n * 4, cudaMemcpyHostToDevice, st.s));
        }
        std::unique_ptr<lapper_t> l(new lapper_t());
        l->device = device;
        l->build_flags = flags;
        build_index_from_unsorted(l->ix, d_s.p, d_e.p, n, flags, st.s);
        l->next_input_id = n;
        d_s.release();
        d_e.release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        *out = l.release();
    });
}

int lapper_build_u32(const uint32_t *starts, const uint32_t *stops, uint64_t n, int device, lapper_t **out) {
    return lapper_build_u32_ex(starts, stops, n, device, LAPPER_BUILD_DEFAULT, out);
}

int lapper_from_sorted_dev(const uint32_t *d_starts, const uint32_t *d_stops_by_iv, const uint32_t *d_stops_sorted,
                           const uint32_t *d_perm, uint64_t n, int overlaps_merged, int device, uint32_t flags,
                           void *stream, lapper_t **out) {
    return guarded([&] {
        LAPPER_REQUIRE(out, "null output handle");
        *out = nullptr;
        LAPPER_REQUIRE(n == 0 || (d_starts && d_stops_by_iv && d_stops_sorted && d_perm), "null index arrays");
        LAPPER_REQUIRE(n < 0xFFFFFFF0ull, "n must be < 2^32 - 16");
        require_device(device);
        DeviceGuard dg(device);
        cudaStream_t s = (cudaStream_t)stream;
        std::unique_ptr<lapper_t> l(new lapper_t());
        l->device = device;
        l->build_flags = flags;
        l->overlaps_merged = overlaps_merged != 0;
        l->next_input_id = n;
        DeviceIndex &ix = l->ix;
        ix.n = n;
        try {
            if (n) {
                const uint32_t *src[4] = {d_starts, d_stops_by_iv, d_stops_sorted, d_perm};
                uint32_t **dst[4] = {&ix.starts, &ix.stops_by_iv, &ix.stops_sorted, &ix.perm};
                for (int i = 0; i < 4; i++) {
                    LAPPER_CUDA_CHECK(cudaMalloc((void **)dst[i], n * 4));
                    ix.bytes += n * 4;
                    LAPPER_CUDA_CHECK(cudaMemcpyAsync(*dst[i], src[i], n * 4, cudaMemcpyDefault, s));
                }
            }
            finish_index_from_sorted(ix, flags, s);
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        } catch (...) {
            free_index(ix);
            throw;
        }
        *out = l.release();
    });
}

void lapper_free(lapper_t *l) {
    if (!l) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(l->device);
    free_index(l->ix);
    if (prev >= 0) cudaSetDevice(prev);
    delete l;
}

uint64_t lapper_len(const lapper_t *l) { return l ? l->ix.n : 0; }
uint32_t lapper_max_len(const lapper_t *l) { return l ? l->ix.max_len : 0; }
int lapper_overlaps_merged(const lapper_t *l) { return l && l->overlaps_merged; }
int lapper_has_inverted(const lapper_t *l) { return l && l->ix.has_inverted; }
int lapper_device(const lapper_t *l) { return l ? l->device : -1; }
uint32_t lapper_bucket_shift(const lapper_t *l) { return l ? l->ix.shift : 0xFFFFFFFFu; }
uint64_t lapper_index_bytes(const lapper_t *l) {
    return l ? l->ix.bytes : 0;
}

uint64_t lapper_query_kernel_launches(void) { return lap::g_query_launches.load(); }

/* -DLAPPER_DEBUG_CHECKS builds: number of bounds checks the query kernels have seen violated on the current device
 * (synchronises it); -1 in normal builds */
long long lapper_debug_violations(void) { return lap::debug_violations(); }

int lapper_packed_table_info(const lapper_t *l, uint32_t *width_out, uint64_t *sectors_out, uint64_t *overflowed_out) {
    return guarded([&] {
        check_handle(l);
        const bool built = l->ix.dense != nullptr;
        if (width_out) *width_out = built ? l->ix.dense_w : 0u;
        if (sectors_out) *sectors_out = built ? l->ix.dense_nb + 1 : 0u;
        if (overflowed_out) *overflowed_out = built ? l->ix.dense_overflow : 0u;
    });
}

int lapper_get_intervals(const lapper_t *l, uint32_t *starts_out, uint32_t *stops_out, uint32_t *perm_out) {
    return guarded([&] {
        check_handle(l);
        const uint64_t n = l->ix.n;
        if (n == 0) return;
        DeviceGuard dg(l->device);
        if (starts_out) LAPPER_CUDA_CHECK(cudaMemcpy(starts_out, l->ix.starts, n * 4, cudaMemcpyDeviceToHost));
        if (stops_out) LAPPER_CUDA_CHECK(cudaMemcpy(stops_out, l->ix.stops_by_iv, n * 4, cudaMemcpyDeviceToHost));
        if (perm_out) LAPPER_CUDA_CHECK(cudaMemcpy(perm_out, l->ix.perm, n * 4, cudaMemcpyDeviceToHost));
    });
}

int lapper_get_starts(const lapper_t *l, uint32_t *out) { return lapper_get_intervals(l, out, nullptr, nullptr); }

int lapper_get_stops(const lapper_t *l, uint32_t *out) {
    return guarded([&] {
        check_handle(l);
        if (l->ix.n == 0) return;
        LAPPER_REQUIRE(out, "null pointer");
        DeviceGuard dg(l->device);
        LAPPER_CUDA_CHECK(cudaMemcpy(out, l->ix.stops_sorted, l->ix.n * 4, cudaMemcpyDeviceToHost));
    });
}

int lapper_get_dev_view(const lapper_t *l, lapper_dev_view *out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(out, "null pointer");
        out->starts = l->ix.starts;
        out->stops_by_iv = l->ix.stops_by_iv;
        out->stops_sorted = l->ix.stops_sorted;
        out->perm = l->ix.perm;
        out->prefix_max = l->ix.prefix_max;
        out->n = l->ix.n;
    });
}

int lapper_count_batch_u32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                           uint64_t *counts_out) {
    return guarded([&] { count_host<uint64_t>(l, q_start, q_stop, nq, counts_out); });
}

int lapper_count_batch_u32_c32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                               uint32_t *counts_out) {
    return guarded([&] { count_host<uint32_t>(l, q_start, q_stop, nq, counts_out); });
}

int lapper_count_batch_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                               uint32_t *d_counts_out, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_counts_out), "count_batch: null pointer");
        DeviceGuard dg(l->device);
        launch_count(l->ix, d_q_start, d_q_stop, nq, d_counts_out, nullptr, (cudaStream_t)stream);
    });
}

int lapper_count_batch_u32_dev64(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                                 uint64_t *d_counts_out, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_counts_out), "count_batch: null pointer");
        DeviceGuard dg(l->device);
        launch_count(l->ix, d_q_start, d_q_stop, nq, nullptr, d_counts_out, (cudaStream_t)stream);
    });
}

int lapper_find_batch_u32_devq(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                               void *stream, lapper_result_t **out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(out, "null output");
        *out = nullptr;
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop), "find_batch: null pointer");
        DeviceGuard dg(l->device);
        *out = find_on_device(l, d_q_start, d_q_stop, nq, (cudaStream_t)stream);
    });
}

int lapper_find_batch_u32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                          lapper_result_t **out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(out, "null output");
        *out = nullptr;
        LAPPER_REQUIRE(nq == 0 || (q_start && q_stop), "find_batch: null pointer");
        DeviceGuard dg(l->device);
        StreamOwner st;
        Scratch<uint32_t> d_qs(nq, st.s), d_qe(nq, st.s);
        if (nq) {
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qs.p, q_start, nq * 4, cudaMemcpyHostToDevice, st.s));
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qe.p, q_stop, nq * 4, cudaMemcpyHostToDevice, st.s));
        }
        *out = find_on_device(l, d_qs.p, d_qe.p, nq, st.s);
        d_qs.release();
        d_qe.release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

uint64_t lapper_result_nq(const lapper_result_t *r) { return r ? r->nq : 0; }
uint64_t lapper_result_total(const lapper_result_t *r) { return r ? r->total : 0; }

int lapper_result_copy(const lapper_result_t *r, uint64_t *offsets_out, uint32_t *indices_out) {
    return guarded([&] {
        LAPPER_REQUIRE(r, "null result");
        DeviceGuard dg(r->device);
        if (offsets_out)
            LAPPER_CUDA_CHECK(cudaMemcpy(offsets_out, r->d_offsets, (r->nq + 1) * 8, cudaMemcpyDeviceToHost));
        if (indices_out && r->total)
            LAPPER_CUDA_CHECK(cudaMemcpy(indices_out, r->d_indices, r->total * 4, cudaMemcpyDeviceToHost));
    });
}

int lapper_result_dev(const lapper_result_t *r, const uint64_t **d_offsets, const uint32_t **d_indices) {
    return guarded([&] {
        LAPPER_REQUIRE(r, "null result");
        if (d_offsets) *d_offsets = r->d_offsets;
        if (d_indices) *d_indices = r->d_indices;
    });
}

void lapper_result_free(lapper_result_t *r) {
    if (!r) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(r->device);
    cudaFree(r->d_offsets);
    cudaFree(r->d_indices);
    if (prev >= 0) cudaSetDevice(prev);
    delete r;
}

int lapper_find_batch_offsets_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                      uint64_t nq, uint64_t *d_offsets, uint64_t *total_out, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(d_offsets && (nq == 0 || (d_q_start && d_q_stop)), "find_batch_offsets: null pointer");
        DeviceGuard dg(l->device);
        const uint64_t t = launch_find_offsets(l->ix, d_q_start, d_q_stop, nq, d_offsets, (cudaStream_t)stream);
        if (total_out) *total_out = t;
    });
}

int lapper_find_batch_fill_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint64_t nq, const uint64_t *d_offsets, uint32_t *d_indices, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_offsets), "find_batch_fill: null pointer");
        DeviceGuard dg(l->device);
        launch_find_fill(l->ix, d_q_start, d_q_stop, nq, d_offsets, d_indices, (cudaStream_t)stream);
    });
}

int lapper_seek_cursor(const lapper_t *l, uint32_t start, uint64_t *cursor) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(cursor, "null cursor");
        if (l->ix.n == 0) {
            // lower_bound over an empty vector is 0 and the while-loop never runs (src/lib.rs:658-671)
            if (*cursor == 0) *cursor = 0;
            return;
        }
        DeviceGuard dg(l->device);
        StreamOwner st;
        Scratch<uint64_t> d_c(1, st.s);
        seek_cursor_kernel<<<1, 1, 0, st.s>>>(l->ix.starts, (uint32_t)l->ix.n, l->ix.max_len, start, *cursor, d_c.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(cursor, d_c.p, 8, cudaMemcpyDeviceToHost, st.s));
        d_c.release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper_insert_u32(lapper_t *l, uint32_t start, uint32_t stop, uint64_t *input_position_out) {
    return guarded([&] {
        check_handle(l);
        DeviceGuard dg(l->device);
        StreamOwner st;
        DeviceIndex &ix = l->ix;
        const uint64_t n = ix.n;
        LAPPER_REQUIRE(n + 1 < 0xFFFFFFF0ull, "lapper_insert: too many intervals");
        uint32_t pos[2] = {0, 0};
        if (n) {
            Scratch<uint32_t> d_pos(2, st.s);
            insert_positions_kernel<<<1, 1, 0, st.s>>>(ix.starts, ix.stops_by_iv, ix.stops_sorted, (uint32_t)n, start, stop, d_pos.p);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(pos, d_pos.p, 8, cudaMemcpyDeviceToHost, st.s));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        }
        const uint64_t id = l->next_input_id;
        uint32_t *arr[4] = {nullptr, nullptr, nullptr, nullptr};
        try {
            for (auto &a : arr) LAPPER_CUDA_CHECK(cudaMalloc((void **)&a, (n + 1) * 4));
            insert_into(arr[0], ix.starts, n, pos[0], start, st.s);
            insert_into(arr[1], ix.stops_by_iv, n, pos[0], stop, st.s);
            insert_into(arr[2], ix.stops_sorted, n, pos[1], stop, st.s);
            insert_into(arr[3], ix.perm, n, pos[0], (uint32_t)id, st.s);
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));  // the host-side scalars above are read by the copies
        } catch (...) {
            for (auto a : arr) cudaFree(a);
            throw;
        }
        free_index(ix);
        ix.n = n + 1;
        ix.starts = arr[0];
        ix.stops_by_iv = arr[1];
        ix.stops_sorted = arr[2];
        ix.perm = arr[3];
        ix.bytes = 16 * (n + 1);
        finish_index_from_sorted(ix, l->build_flags, st.s);  // max_len (src/lib.rs:264-267) falls out of the rebuild
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        l->next_input_id = id + 1;
        l->cov_is_some = false;       // src/lib.rs:271
        l->overlaps_merged = false;   // src/lib.rs:272
        if (input_position_out) *input_position_out = id;
    });
}

int lapper_merge_overlaps(lapper_t *l) {
    return guarded([&] {
        check_handle(l);
        if (l->ix.n == 0) return;  // src/lib.rs:366: nothing happens without a first interval
        need_valid_intervals(l, "merge_overlaps");
        DeviceGuard dg(l->device);
        StreamOwner st;
        uint32_t *s = nullptr, *e = nullptr, *p = nullptr;
        const uint64_t m = compute_merged(l->ix, &s, &e, &p, st.s);
        try {
            replace_index_with_sorted(l, s, e, p, m, st.s);
        } catch (...) {
            throw;
        }
        l->overlaps_merged = true;  // src/lib.rs:382; the cached cov is left alone like the reference
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper_cov(const lapper_t *l, uint32_t *cov_out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(cov_out, "null pointer");
        if (l->cov_is_some) {  // src/lib.rs:312-315
            *cov_out = l->cov;
            return;
        }
        need_valid_intervals(l, "cov");
        DeviceGuard dg(l->device);
        StreamOwner st;
        *cov_out = compute_cov(l->ix, st.s);
    });
}

/* the cached fields of src/lib.rs:119-122 (`cov: Option<I>`, `overlaps_merged`), for (de)serialisation */
int lapper_get_cached_state(const lapper_t *l, int *cov_is_some_out, uint32_t *cov_out, int *overlaps_merged_out) {
    return guarded([&] {
        check_handle(l);
        if (cov_is_some_out) *cov_is_some_out = l->cov_is_some ? 1 : 0;
        if (cov_out) *cov_out = l->cov_is_some ? l->cov : 0u;
        if (overlaps_merged_out) *overlaps_merged_out = l->overlaps_merged ? 1 : 0;
    });
}

int lapper_restore_cached_state(lapper_t *l, int cov_is_some, uint32_t cov, int overlaps_merged) {
    return guarded([&] {
        check_handle(l);
        l->cov_is_some = cov_is_some != 0;
        l->cov = cov_is_some ? cov : 0u;
        l->overlaps_merged = overlaps_merged != 0;
    });
}

int lapper_set_cov(lapper_t *l, uint32_t *cov_out) {
    return guarded([&] {
        check_handle(l);
        need_valid_intervals(l, "set_cov");
        DeviceGuard dg(l->device);
        StreamOwner st;
        const uint32_t c = compute_cov(l->ix, st.s);  // src/lib.rs:320-324
        l->cov = c;
        l->cov_is_some = true;
        if (cov_out) *cov_out = c;
    });
}

int lapper_union_and_intersect(const lapper_t *a, const lapper_t *b, uint32_t *union_out, uint32_t *intersect_out) {
    return guarded([&] {
        check_handle(a);
        check_handle(b);
        LAPPER_REQUIRE(a->device == b->device, "union_and_intersect: both lappers must live on the same device");
        need_valid_intervals(a, "union_and_intersect");
        need_valid_intervals(b, "union_and_intersect");
        DeviceGuard dg(a->device);
        StreamOwner st;
        const uint32_t x = compute_intersect_measure(a->ix, b->ix, st.s);
        const uint32_t ca = a->cov_is_some ? a->cov : compute_cov(a->ix, st.s);
        const uint32_t cb = b->cov_is_some ? b->cov : compute_cov(b->ix, st.s);
        if (union_out) *union_out = ca + cb - x;  // src/lib.rs:532 / :542
        if (intersect_out) *intersect_out = x;
    });
}

int lapper_depth(const lapper_t *l, uint64_t *n_out, uint32_t **starts_out, uint32_t **stops_out, uint32_t **depth_out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(n_out && starts_out && stops_out && depth_out, "lapper_depth: null pointer");
        *n_out = 0;
        *starts_out = *stops_out = *depth_out = nullptr;
        LAPPER_REQUIRE(l->ix.n > 0, "depth() on an empty lapper: the reference panics here (src/lib.rs:744)");
        need_valid_intervals(l, "depth");
        DeviceGuard dg(l->device);
        StreamOwner st;
        uint32_t *d[3] = {nullptr, nullptr, nullptr};
        const uint64_t nr = compute_depth(l->ix, &d[0], &d[1], &d[2], st.s);
        uint32_t *h[3] = {nullptr, nullptr, nullptr};
        try {
            for (int i = 0; i < 3; i++) {
                h[i] = (uint32_t *)malloc(std::max<uint64_t>(nr, 1) * 4);
                if (!h[i]) throw std::bad_alloc();
                if (nr) LAPPER_CUDA_CHECK(cudaMemcpy(h[i], d[i], nr * 4, cudaMemcpyDeviceToHost));
            }
        } catch (...) {
            for (int i = 0; i < 3; i++) {
                free(h[i]);
                cudaFree(d[i]);
            }
            throw;
        }
        for (int i = 0; i < 3; i++) cudaFree(d[i]);
        *n_out = nr;
        *starts_out = h[0];
        *stops_out = h[1];
        *depth_out = h[2];
    });
}

void lapper_host_free(void *p) { free(p); }

int lapper_host_alloc_pinned(uint64_t bytes, void **out) {
    return guarded([&] {
        LAPPER_REQUIRE(out, "null pointer");
        *out = nullptr;
        LAPPER_CUDA_CHECK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    });
}

void lapper_host_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------- multi-GPU group
int lapper_replicate(const lapper_t *l, const int *devices, int g, lapper_group_t **out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(out && devices && g > 0, "lapper_replicate: bad arguments");
        *out = nullptr;
        std::unique_ptr<lapper_group_t> grp(new lapper_group_t());
        try {
            for (int i = 0; i < g; i++) {
                require_device(devices[i]);
                DeviceGuard dg(devices[i]);
                StreamOwner st;
                lapper_t *rep = nullptr;
                // cudaMemcpyDefault with unified addressing: a peer copy over NVLink when the source
                // lives on another GPU
                int rc = lapper_from_sorted_dev(l->ix.starts, l->ix.stops_by_iv, l->ix.stops_sorted, l->ix.perm, l->ix.n,
                                                l->overlaps_merged, devices[i], l->build_flags, st.s, &rep);
                if (rc != LAPPER_OK) throw Error(rc, g_last_error);
                rep->cov_is_some = l->cov_is_some;
                rep->cov = l->cov;
                grp->replicas.push_back(rep);
            }
        } catch (...) {
            for (lapper_t *r : grp->replicas) lapper_free(r);
            throw;
        }
        *out = grp.release();
    });
}

void lapper_group_free(lapper_group_t *grp) {
    if (!grp) return;
    for (lapper_t *r : grp->replicas) lapper_free(r);
    delete grp;
}

int lapper_group_size(const lapper_group_t *grp) { return grp ? (int)grp->replicas.size() : 0; }

const lapper_t *lapper_group_replica(const lapper_group_t *grp, int i) {
    if (!grp || i < 0 || i >= (int)grp->replicas.size()) return nullptr;
    return grp->replicas[i];
}

int lapper_group_count_batch_u32(const lapper_group_t *grp, const uint32_t *q_start, const uint32_t *q_stop,
                                 uint64_t nq, uint64_t *counts_out) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty(), "null group");
        if (nq == 0) return;
        LAPPER_REQUIRE(q_start && q_stop && counts_out, "group_count_batch: null pointer");
        const uint64_t G = grp->replicas.size();
        struct Shard {
            std::unique_ptr<StreamOwner> st;
            Scratch<uint32_t> qs, qe;
            Scratch<uint64_t> out;
        };
        std::vector<Shard> sh(G);
        // enqueue every shard asynchronously (device g owns [g*nq/G, (g+1)*nq/G)), then wait for all
        for (uint64_t g = 0; g < G; g++) {
            const uint64_t lo = nq * g / G, hi = nq * (g + 1) / G, m = hi - lo;
            if (m == 0) continue;
            const lapper_t *rep = grp->replicas[g];
            DeviceGuard dg(rep->device);
            sh[g].st.reset(new StreamOwner());
            cudaStream_t s = sh[g].st->s;
            sh[g].qs.alloc(m, s);
            sh[g].qe.alloc(m, s);
            sh[g].out.alloc(m, s);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(sh[g].qs.p, q_start + lo, m * 4, cudaMemcpyHostToDevice, s));
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(sh[g].qe.p, q_stop + lo, m * 4, cudaMemcpyHostToDevice, s));
            launch_count(rep->ix, sh[g].qs.p, sh[g].qe.p, m, nullptr, sh[g].out.p, s);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(counts_out + lo, sh[g].out.p, m * 8, cudaMemcpyDeviceToHost, s));
        }
        for (uint64_t g = 0; g < G; g++) {
            if (!sh[g].st) continue;
            DeviceGuard dg(grp->replicas[g]->device);
            sh[g].qs.release();
            sh[g].qe.release();
            sh[g].out.release();
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(sh[g].st->s));
        }
    });
}

int lapper_group_find_batch_u32(const lapper_group_t *grp, const uint32_t *q_start, const uint32_t *q_stop,
                                uint64_t nq, uint64_t *offsets_out, uint32_t **indices_out, uint64_t *total_out) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty(), "null group");
        LAPPER_REQUIRE(offsets_out && indices_out && total_out, "group_find_batch: null pointer");
        *indices_out = nullptr;
        *total_out = 0;
        const uint64_t G = grp->replicas.size();
        std::vector<lapper_result_t *> res(G, nullptr);
        auto cleanup = [&] {
            for (auto *r : res) lapper_result_free(r);
        };
        try {
            for (uint64_t g = 0; g < G; g++) {
                const uint64_t lo = nq * g / G, hi = nq * (g + 1) / G;
                int rc = lapper_find_batch_u32(grp->replicas[g], q_start + lo, q_stop + lo, hi - lo, &res[g]);
                if (rc != LAPPER_OK) throw Error(rc, g_last_error);
            }
            // rebase each shard's CSR by the totals of the shards before it
            uint64_t total = 0;
            for (auto *r : res) total += r->total;
            uint32_t *idx = (uint32_t *)malloc(std::max<uint64_t>(total, 1) * 4);
            if (!idx) throw std::bad_alloc();
            uint64_t base = 0;
            for (uint64_t g = 0; g < G; g++) {
                const uint64_t lo = nq * g / G, m = res[g]->nq;
                int rc = lapper_result_copy(res[g], offsets_out + lo, idx + base);
                if (rc != LAPPER_OK) {
                    free(idx);
                    throw Error(rc, g_last_error);
                }
                if (base)
                    for (uint64_t j = 0; j <= m; j++) offsets_out[lo + j] += base;
                base += res[g]->total;
            }
            offsets_out[nq] = total;
            *indices_out = idx;
            *total_out = total;
        } catch (...) {
            cleanup();
            throw;
        }
        cleanup();
    });
}

}  // extern "C"
