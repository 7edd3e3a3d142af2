// api.cu -- the extern "C" surface declared in include/lapper_b200.h.
// Status codes only; no exception or abort crosses the ABI.  No CPU fallback anywhere: every entry
// point that computes runs the CUDA kernels of build.cu / query.cu / setops.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "abi_common.cuh"
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

uint32_t sm_count() {
    static std::atomic<uint32_t> cached[64];
    int dev = 0;
    LAPPER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) throw Error(LAPPER_ERR_INVALID_ARG, "device index out of range");
    uint32_t v = cached[dev].load(std::memory_order_relaxed);
    if (!v) {
        int n = 0;
        LAPPER_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        v = (uint32_t)std::max(n, 1);
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
}  // namespace lap

namespace {

thread_local std::string g_last_error;
void require_device(int device);
}  // namespace
namespace lap {
std::string &last_error_slot() { return g_last_error; }
void require_device_index(int device) { require_device(device); }
}  // namespace lap
namespace {

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

using StreamOwner = OwnedStream;  // abi_common.cuh: per-thread cache of the streams calls run on

void check_handle(const lapper_t *l) { LAPPER_REQUIRE(l != nullptr, "null lapper handle"); }

// count / find need the query tables; after merge_overlaps / insert they are built here, on first use.  Concurrent
// const calls on one handle are allowed (include/lapper_b200.h), hence the lock.
void ensure_tables(const lapper_t *cl, cudaStream_t stream) {
    if (cl->tables_flag.load(std::memory_order_acquire)) return;
    lapper_t *l = const_cast<lapper_t *>(cl);
    std::lock_guard<std::mutex> lock(l->tables_mu);
    if (!l->ix.tables_ready) build_query_tables(l->ix, l->build_flags, stream);  // synchronises `stream` at its end
    l->tables_flag.store(true, std::memory_order_release);
}

void need_valid_intervals(const lapper_t *l, const char *what) {
    if (l->ix.has_inverted)
        throw Error(LAPPER_ERR_PRECONDITION,
                    std::string(what) + ": the set holds an interval with stop < start; the closed form used here "
                                        "(prefix max of stops, SURVEY Appendix A.5-A.7) needs start <= stop for every "
                                        "interval -- documented deviation, see include/lapper_b200.h");
}

// A DeviceIndex under construction: its arrays go back to the pool unless it is committed into a handle.
struct IndexBuilder {
    DeviceIndex ix;
    cudaStream_t stream;
    bool committed = false;
    explicit IndexBuilder(cudaStream_t s) : stream(s) {}
    ~IndexBuilder() {
        if (!committed) free_index(ix, stream);
    }
    // replaces l's index (the old arrays are returned to the pool in stream order)
    void commit(lapper_t *l) {
        ix.query_len_hint = l->ix.query_len_hint;  // (the announced query length outlives merge_overlaps / insert)
        free_index(l->ix, stream);
        l->ix = ix;
        l->tables_flag.store(ix.tables_ready, std::memory_order_release);
        committed = true;
    }
};

}  // namespace
namespace lap {
// ---- pageable callers.  A Vec<u32> / numpy array / std::vector is pageable memory: cudaMemcpyAsync from it runs at
// 11 GB/s in and 22 GB/s out on the bench boxes (the driver stages it through its own small pinned buffers, one
// thread), against 55 GB/s for pinned memory (tools/mb_pageable.py).  For large batches the library does the staging
// itself: pinned staging buffers (cached for the life of the process), filled and drained by a few host threads
// (memcpy at ~15 GB/s each), with the copies of chunk k+1 and k-1 running beside the kernel of chunk k.
bool host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int host_copy_threads() {
    static const int n = [] {
        const char *e = std::getenv("LAPPER_HOST_COPY_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        return std::min(std::max(e ? atoi(e) : 4, 0), std::max(hw, 1));  // 0: leave pageable buffers to the driver
    }();
    return n;
}

// memcpy in host_copy_threads() slices (the calling thread takes the last one)
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    const int nt = bytes < (4u << 20) ? 1 : host_copy_threads();
    if (nt <= 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t slice = ((bytes / nt) + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    size_t off = 0;
    for (int i = 0; i < nt - 1 && off + slice < bytes; i++, off += slice)
        th.emplace_back([=] { std::memcpy((char *)dst + off, (const char *)src + off, slice); });
    std::memcpy((char *)dst + off, (const char *)src + off, bytes - off);
    for (auto &x : th) x.join();
}

// pinned staging buffers, kept for the life of the process (cudaHostAlloc pins pages at a few GB/s: not per call)
struct PinnedCache {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_list;
    void *get(size_t bytes, size_t &got) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].second >= bytes) {
                    void *p = free_list[i].first;
                    got = free_list[i].second;
                    free_list.erase(free_list.begin() + i);
                    return p;
                }
        }
        void *p = nullptr;
        LAPPER_CUDA_CHECK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
        got = bytes;
        return p;
    }
    void put(void *p, size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        free_list.emplace_back(p, bytes);
    }
};
PinnedCache &pinned_cache() {
    static PinnedCache *c = new PinnedCache();  // (never destroyed: the driver may be gone at static destruction time)
    return *c;
}
PinnedBuf::PinnedBuf(size_t n) { p = pinned_cache().get(n, bytes); }
PinnedBuf::~PinnedBuf() {
    if (p) pinned_cache().put(p, bytes);
}

// One-directional copies between device memory and PAGEABLE host memory, staged through two pinned buffers: while the
// copy threads move piece k between the caller's memory and one buffer, the DMA engine moves piece k+1 through the
// other.  Synchronous (returns when the data is in place); small or pinned transfers take plain cudaMemcpy.
constexpr size_t kStagePiece = 32u << 20;
void copy_to_host(void *h_dst, const void *d_src, size_t bytes) {
    if (bytes < 2 * kStagePiece || host_copy_threads() == 0 || host_pointer_is_pinned(h_dst)) {
        LAPPER_CUDA_CHECK(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
        return;
    }
    StreamOwner st;
    PinnedBuf buf[2] = {PinnedBuf(kStagePiece), PinnedBuf(kStagePiece)};
    const size_t npieces = (bytes + kStagePiece - 1) / kStagePiece;
    const auto piece_bytes = [&](size_t k) { return std::min(kStagePiece, bytes - k * kStagePiece); };
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(buf[0].p, d_src, piece_bytes(0), cudaMemcpyDeviceToHost, st.s));
    for (size_t k = 0; k < npieces; k++) {
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));  // piece k has arrived in buf[k & 1]
        if (k + 1 < npieces)
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(buf[(k + 1) & 1].p, (const char *)d_src + (k + 1) * kStagePiece, piece_bytes(k + 1),
                                              cudaMemcpyDeviceToHost, st.s));
        parallel_memcpy((char *)h_dst + k * kStagePiece, buf[k & 1].p, piece_bytes(k));
    }
}
void copy_to_device(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream) {
    if (bytes < 2 * kStagePiece || host_copy_threads() == 0 || host_pointer_is_pinned(h_src)) {
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        return;
    }
    PinnedBuf buf[2] = {PinnedBuf(kStagePiece), PinnedBuf(kStagePiece)};
    cudaEvent_t done[2];
    LAPPER_CUDA_CHECK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    LAPPER_CUDA_CHECK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    const size_t npieces = (bytes + kStagePiece - 1) / kStagePiece;
    cudaError_t err = cudaSuccess;
    for (size_t k = 0; k < npieces && err == cudaSuccess; k++) {
        const size_t b = std::min(kStagePiece, bytes - k * kStagePiece);
        if (k >= 2) err = cudaEventSynchronize(done[k & 1]);  // the buffer's previous piece has left it
        if (err != cudaSuccess) break;
        parallel_memcpy(buf[k & 1].p, (const char *)h_src + k * kStagePiece, b);
        err = cudaMemcpyAsync((char *)d_dst + k * kStagePiece, buf[k & 1].p, b, cudaMemcpyHostToDevice, stream);
        if (err == cudaSuccess) err = cudaEventRecord(done[k & 1], stream);
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);  // the staging buffers go back to the cache on return
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    LAPPER_CUDA_CHECK(err);
}

}  // namespace lap
namespace {
constexpr uint64_t kStagedMinQueries = 4ull << 20;  // below this the driver's own pageable path is as good

// count_host for pageable buffers: chunks of 4 M queries, three slots; per chunk the host threads copy the queries into
// the slot's pinned buffers, the stream does H2D | count | D2H into the slot's pinned output, and the counts of the
// chunk that used the slot before are copied out to the caller first.
template <typename OutT>
void count_host_staged(const lapper_t *l, const uint32_t *qs, const uint32_t *qe, uint64_t nq, OutT *out) {
    constexpr int kSlots = 3;
    const uint64_t chunk = std::min<uint64_t>(nq, 1ull << 22);
    const int used = (int)std::min<uint64_t>(kSlots, div_up(nq, chunk));
    StreamOwner st[kSlots];
    ensure_tables(l, st[0].s);
    Scratch<uint32_t> d_qs[kSlots], d_qe[kSlots];
    Scratch<OutT> d_out[kSlots];
    std::vector<PinnedBuf> h_qs, h_qe, h_out;
    uint64_t pending_lo[kSlots] = {}, pending_m[kSlots] = {};  // the chunk whose counts sit in the slot's pinned output
    for (int s = 0; s < used; s++) {
        d_qs[s].alloc(chunk, st[s].s), d_qe[s].alloc(chunk, st[s].s), d_out[s].alloc(chunk, st[s].s);
        h_qs.emplace_back(chunk * 4), h_qe.emplace_back(chunk * 4), h_out.emplace_back(chunk * sizeof(OutT));
    }
    const auto drain = [&](int s) {
        if (!pending_m[s]) return;
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st[s].s));
        parallel_memcpy(out + pending_lo[s], h_out[s].p, pending_m[s] * sizeof(OutT));
        pending_m[s] = 0;
    };
    uint64_t done = 0;
    for (int i = 0; done < nq; i++) {
        const int s = i % kSlots;
        const uint64_t m = std::min(chunk, nq - done);
        drain(s);  // (also: the slot's H2D of its previous chunk has finished, its pinned inputs are free)
        parallel_memcpy(h_qs[s].p, qs + done, m * 4);
        parallel_memcpy(h_qe[s].p, qe + done, m * 4);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qs[s].p, h_qs[s].p, m * 4, cudaMemcpyHostToDevice, st[s].s));
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qe[s].p, h_qe[s].p, m * 4, cudaMemcpyHostToDevice, st[s].s));
        if (sizeof(OutT) == 8)
            launch_count(l->ix, d_qs[s].p, d_qe[s].p, m, nullptr, reinterpret_cast<uint64_t *>(d_out[s].p), st[s].s);
        else
            launch_count(l->ix, d_qs[s].p, d_qe[s].p, m, reinterpret_cast<uint32_t *>(d_out[s].p), nullptr, st[s].s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(h_out[s].p, d_out[s].p, m * sizeof(OutT), cudaMemcpyDeviceToHost, st[s].s));
        pending_lo[s] = done, pending_m[s] = m;
        done += m;
    }
    for (int k = 0; k < kSlots; k++) drain((int)((div_up(nq, chunk) + k) % kSlots));  // oldest chunk first; empty slots are no-ops
    for (int s = 0; s < used; s++) {
        d_qs[s].release(), d_qe[s].release(), d_out[s].release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st[s].s));
    }
}

// chunked host <-> device pipeline for count: three slots, each its own stream, so the H2D of chunk
// i+1, the kernel of chunk i and the D2H of chunk i-1 overlap (PCIe is full duplex).
template <typename OutT>
void count_host(const lapper_t *l, const uint32_t *qs, const uint32_t *qe, uint64_t nq, OutT *out) {
    check_handle(l);
    if (nq == 0) return;
    LAPPER_REQUIRE(qs && qe && out, "count_batch: null pointer");
    DeviceGuard dg(l->device);
    if (nq >= kStagedMinQueries && host_copy_threads() > 0 && !(host_pointer_is_pinned(qs) && host_pointer_is_pinned(qe) && host_pointer_is_pinned(out))) {
        count_host_staged<OutT>(l, qs, qe, nq, out);
        return;
    }
    constexpr int kSlots = 3;
    // 8 M queries per chunk: measured best of 2^20 .. 2^25 (20.0 / 18.0 / 17.1 / 16.8 / 17.0 / 18.0 ms per 100 M queries)
    const uint64_t chunk = std::min<uint64_t>(nq, 1ull << 23);
    StreamOwner st[kSlots];
    ensure_tables(l, st[0].s);
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

void free_result_arrays(lapper_result_t *r, cudaStream_t stream) {
    pool_free(r->d_offsets, stream);
    pool_free(r->d_indices, stream);
    r->d_offsets = nullptr;
    r->d_indices = nullptr;
}

lapper_result_t *find_on_device(const lapper_t *l, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                                cudaStream_t stream) {
    ensure_tables(l, stream);
    std::unique_ptr<lapper_result_t> r(new lapper_result_t());
    r->device = l->device;
    r->nq = nq;
    try {
        uint64_t bytes = 0;
        r->d_offsets = pool_alloc<uint64_t>(nq + 1, bytes, stream);
        // Both phases in one enqueue when the size of the answer can be guessed: from the handle's previous batch
        // (+ 1/8), or 16 positions per query for a first batch that is small (<= 64 Mi positions).  A wrong guess costs
        // nothing but the buffer: the offsets are complete, and the emit runs again into a buffer of the right size.
        const uint64_t ratio = l->find_hits_per_query_x256.load(std::memory_order_relaxed);
        uint64_t guess = 0;
        if (ratio)
            guess = (uint64_t)(((unsigned __int128)nq * ratio) >> 8) / 8 * 9 + 4096;
        else if (nq <= (64ull << 20) / 16)
            guess = nq * 16;
        if (guess && nq) {
            r->d_indices = pool_alloc<uint32_t>(guess, bytes, stream);
            r->total = launch_find(l->ix, d_qs, d_qe, nq, r->d_offsets, r->d_indices, guess, stream);
            if (r->total > guess) {
                pool_free(r->d_indices, stream);
                r->d_indices = nullptr;
                r->d_indices = pool_alloc<uint32_t>(r->total, bytes, stream);
                launch_find_fill(l->ix, d_qs, d_qe, nq, r->d_offsets, r->d_indices, stream, r->total);
            }
        } else {
            r->total = launch_find_offsets(l->ix, d_qs, d_qe, nq, r->d_offsets, stream);
            r->d_indices = pool_alloc<uint32_t>(r->total, bytes, stream);
            launch_find_fill(l->ix, d_qs, d_qe, nq, r->d_offsets, r->d_indices, stream, r->total);
        }
        if (nq) l->find_hits_per_query_x256.store(std::max<uint64_t>(div_up(r->total * 256, nq), 1), std::memory_order_relaxed);
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    } catch (...) {
        free_result_arrays(r.get(), stream);
        throw;
    }
    return r.release();
}

// find_batch from host queries straight into host CSR buffers, as a chunk pipeline: while chunk k is counted and
// emitted, chunk k+1's queries are on their way in and chunk k-1's offsets and positions on their way out (three
// slots, one stream each; PCIe is full duplex and the output side -- 8 B of offsets per query + 4 B per hit -- is what
// bounds the call).  A chunk's offsets are computed chunk-local (the emit kernel addresses the chunk's own position
// buffer with them) and rebased by the hits of the chunks before it on their way out.
// Returns the total; positions are written only while they fit `capacity`.
uint64_t find_host(const lapper_t *l, const uint32_t *qs, const uint32_t *qe, uint64_t nq, uint64_t *offsets_out,
                   uint32_t *indices_out, uint64_t capacity, uint64_t offsets_base) {
    check_handle(l);
    DeviceGuard dg(l->device);
    static constexpr int kMaxSlots = 4;
    // 4 M queries per chunk: 32 MB in, ~ 32 MB + 4 B/hit out (experiments: LAPPER_FIND_CHUNK_LOG2, LAPPER_FIND_SLOTS)
    static const uint64_t kChunk = [] {
        const char *e = std::getenv("LAPPER_FIND_CHUNK_LOG2");
        const int k = e ? atoi(e) : 22;
        return 1ull << std::min(std::max(k, 10), 30);
    }();
    static const int kSlots = [] {
        const char *e = std::getenv("LAPPER_FIND_SLOTS");
        return std::min(std::max(e ? atoi(e) : 3, 2), kMaxSlots);
    }();
    StreamOwner st[kMaxSlots];
    ensure_tables(l, st[0].s);
    if (nq == 0) {
        offsets_out[0] = offsets_base;
        return 0;
    }
    // (pageable buffers are left to the driver here: with three streams in flight its own pageable copies come within
    // 5 % of a staged version of this pipeline -- 156 vs 148 ms for 50 M queries / 503 M hits)
    const uint64_t chunk = std::min(nq, kChunk);
    const uint64_t nchunks = div_up(nq, chunk);
    Scratch<uint32_t> d_qs[kMaxSlots], d_qe[kMaxSlots], d_idx[kMaxSlots];
    Scratch<uint64_t> d_off[kMaxSlots], d_goff[kMaxSlots];
    uint64_t idx_cap[kMaxSlots] = {};
    const int used = (int)std::min<uint64_t>(kSlots, nchunks);
    for (int s = 0; s < used; s++) {
        d_qs[s].alloc(chunk, st[s].s);
        d_qe[s].alloc(chunk, st[s].s);
        d_off[s].alloc(chunk + 1, st[s].s);
        d_goff[s].alloc(chunk + 1, st[s].s);
    }
    const auto upload = [&](uint64_t k) {
        const int s = (int)(k % kSlots);
        const uint64_t lo = k * chunk, m = std::min(chunk, nq - lo);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qs[s].p, qs + lo, m * 4, cudaMemcpyHostToDevice, st[s].s));
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(d_qe[s].p, qe + lo, m * 4, cudaMemcpyHostToDevice, st[s].s));
    };
    upload(0);
    uint64_t base = 0;  // hits of the chunks before this one
    for (uint64_t k = 0; k < nchunks; k++) {
        const int s = (int)(k % kSlots);
        const uint64_t lo = k * chunk, m = std::min(chunk, nq - lo);
        if (k + 1 < nchunks) upload(k + 1);  // (its slot's previous chunk drains in stream order first)
        const uint64_t t = launch_find_offsets(l->ix, d_qs[s].p, d_qe[s].p, m, d_off[s].p, st[s].s);  // synchronises st[s]
        const bool last = k + 1 == nchunks;
        launch_rebase_offsets(d_off[s].p, d_goff[s].p, m + (last ? 1 : 0), offsets_base + base, st[s].s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(offsets_out + lo, d_goff[s].p, (m + (last ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, st[s].s));
        if (indices_out && base + t <= capacity && t) {
            // the slot keeps its position buffer while the chunks fit it (+ 1/8 of headroom when it has to grow): the
            // pool sees a handful of allocations per call, of sizes that repeat from call to call
            if (t > idx_cap[s]) {
                idx_cap[s] = t + t / 8;
                d_idx[s].alloc(idx_cap[s], st[s].s);  // (the old buffer goes back to the pool in stream order)
            }
            launch_find_fill(l->ix, d_qs[s].p, d_qe[s].p, m, d_off[s].p, d_idx[s].p, st[s].s, t);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(indices_out + base, d_idx[s].p, t * 4, cudaMemcpyDeviceToHost, st[s].s));
        }
        base += t;
    }
    for (int s = 0; s < used; s++) {
        d_qs[s].release();
        d_qe[s].release();
        d_off[s].release();
        d_goff[s].release();
        d_idx[s].release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st[s].s));
    }
    return base;
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
__global__ void insert_into_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint64_t n, uint64_t pos,
                                   uint32_t value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = i < pos ? src[i] : i == pos ? value : src[i - 1];
}
void insert_into(uint32_t *dst, const uint32_t *src, uint64_t n, uint64_t pos, uint32_t value, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)std::min<uint64_t>(div_up(n + 1, 256), (uint64_t)sm_count() * 8);
    insert_into_kernel<<<grid, 256, 0, stream>>>(dst, src, n, pos, value);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

// copy between devices (or within one); pool memory on either side
void copy_dev_to_dev(void *dst, int dst_dev, const void *src, int src_dev, uint64_t bytes, cudaStream_t stream) {
    if (bytes == 0) return;
    if (dst_dev == src_dev)
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
    else
        LAPPER_CUDA_CHECK(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, stream));
}

int device_of_pointer(const void *p, int fallback) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
    cudaGetLastError();
    return fallback;
}

}  // namespace

struct lapper_group {
    std::vector<lapper_t *> replicas;
};

namespace {

// runs fn(g) for every replica concurrently (one host thread per GPU: each call blocks on its own device) and
// rethrows the first failure
template <typename F>
void for_each_replica(const lapper_group_t *grp, F &&fn) {
    const size_t G = grp->replicas.size();
    std::vector<int> rc(G, LAPPER_OK);
    std::vector<std::string> msg(G);
    const auto body = [&](size_t g) {
        try {
            fn(g);
        } catch (const Error &e) {
            rc[g] = e.code, msg[g] = e.what();
        } catch (const std::exception &e) {
            rc[g] = LAPPER_ERR_CUDA, msg[g] = e.what();
        } catch (...) {
            rc[g] = LAPPER_ERR_CUDA, msg[g] = "unknown error";
        }
    };
    if (G == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(G);
        for (size_t g = 0; g < G; g++) th.emplace_back(body, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++)
        if (rc[g] != LAPPER_OK) throw Error(rc[g], msg[g]);
}

}  // namespace

extern "C" {

int lapper_version(void) { return 200; }

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
        l->tables_flag.store(l->ix.tables_ready);
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
        l->tables_flag.store(l->ix.tables_ready);
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
        IndexBuilder b(s);
        b.ix.n = n;
        if (n) {
            const uint32_t *src[4] = {d_starts, d_stops_by_iv, d_stops_sorted, d_perm};
            uint32_t **dst[4] = {&b.ix.starts, &b.ix.stops_by_iv, &b.ix.stops_sorted, &b.ix.perm};
            for (int i = 0; i < 4; i++) {
                *dst[i] = pool_alloc<uint32_t>(n, b.ix.bytes, s);
                // a peer copy over NVLink when the source lives on another GPU
                copy_dev_to_dev(*dst[i], device, src[i], device_of_pointer(src[i], device), n * 4, s);
            }
        }
        finish_index_core(b.ix, s);
        build_query_tables(b.ix, flags, s);
        // the next inserted interval must not reuse an input position perm already refers to (after merge_overlaps
        // perm holds original positions, which may exceed n)
        l->next_input_id = n ? std::max<uint64_t>(n, (uint64_t)b.ix.stats.max_perm + 1) : 0;
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        b.commit(l.get());
        *out = l.release();
    });
}

void lapper_free(lapper_t *l) {
    if (!l) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(l->device);
    // like cudaFree: work that still uses the index finishes first; the arrays then go back to the library pool
    cudaDeviceSynchronize();
    free_index(l->ix, nullptr);
    if (prev >= 0) cudaSetDevice(prev);
    delete l;
}

uint64_t lapper_len(const lapper_t *l) { return l ? l->ix.n : 0; }
uint32_t lapper_max_len(const lapper_t *l) { return l ? l->ix.max_len : 0; }
int lapper_overlaps_merged(const lapper_t *l) { return l && l->overlaps_merged; }
int lapper_has_inverted(const lapper_t *l) { return l && l->ix.has_inverted; }
int lapper_device(const lapper_t *l) { return l ? l->device : -1; }

/* (the three accessors below describe the query tables: after merge_overlaps / insert they build them if need be) */
static bool tables_for_accessor(const lapper_t *l) {
    if (!l) return false;
    if (l->tables_flag.load(std::memory_order_acquire)) return true;
    return guarded([&] {
               DeviceGuard dg(l->device);
               StreamOwner st;
               ensure_tables(l, st.s);
           }) == LAPPER_OK;
}
uint32_t lapper_bucket_shift(const lapper_t *l) { return tables_for_accessor(l) ? l->ix.shift : 0xFFFFFFFFu; }
uint64_t lapper_index_bytes(const lapper_t *l) { return tables_for_accessor(l) ? l->ix.bytes : 0; }

uint64_t lapper_query_kernel_launches(void) { return lap::g_query_launches.load(); }

/* -DLAPPER_DEBUG_CHECKS builds: number of bounds checks the query kernels have seen violated on the current device
 * (synchronises it); -1 in normal builds */
long long lapper_debug_violations(void) { return lap::debug_violations(); }

int lapper_debug_sorted_run_stats(uint64_t *out2, int reset) {
    return guarded([&] {
        LAPPER_REQUIRE(out2, "null pointer");
        lap::sorted_run_stats(out2, reset != 0);
    });
}

uint64_t lapper_debug_set_sorted_min_queries(uint64_t min_queries) {
    return lap::g_sorted_min_queries.exchange(std::max<uint64_t>(min_queries, 32 * 1024), std::memory_order_relaxed);
}

int lapper_packed_table_info(const lapper_t *l, uint32_t *width_out, uint64_t *sectors_out, uint64_t *overflowed_out) {
    return guarded([&] {
        check_handle(l);
        {
            DeviceGuard dg(l->device);
            StreamOwner st;
            ensure_tables(l, st.s);
        }
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
        ensure_tables(l, (cudaStream_t)stream);
        launch_count(l->ix, d_q_start, d_q_stop, nq, d_counts_out, nullptr, (cudaStream_t)stream);
    });
}

int lapper_count_batch_u32_dev64(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                                 uint64_t *d_counts_out, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_counts_out), "count_batch: null pointer");
        DeviceGuard dg(l->device);
        ensure_tables(l, (cudaStream_t)stream);
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
            copy_to_device(d_qs.p, q_start, nq * 4, st.s);
            copy_to_device(d_qe.p, q_stop, nq * 4, st.s);
        }
        *out = find_on_device(l, d_qs.p, d_qe.p, nq, st.s);
        d_qs.release();
        d_qe.release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper_find_batch_u32_host(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                               uint64_t *offsets_out, uint32_t *indices_out, uint64_t indices_capacity,
                               uint64_t *total_out) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(offsets_out && total_out, "find_batch_host: null output");
        LAPPER_REQUIRE(nq == 0 || (q_start && q_stop), "find_batch_host: null pointer");
        *total_out = 0;
        const uint64_t total = find_host(l, q_start, q_stop, nq, offsets_out, indices_out, indices_capacity, 0);
        *total_out = total;
        if (indices_out && total > indices_capacity)
            throw Error(LAPPER_ERR_CAPACITY, "find_batch_host: " + std::to_string(total) + " positions do not fit the buffer of " +
                                                 std::to_string(indices_capacity) + " (offsets and *total_out are complete)");
    });
}

uint64_t lapper_result_nq(const lapper_result_t *r) { return r ? r->nq : 0; }
uint64_t lapper_result_total(const lapper_result_t *r) { return r ? r->total : 0; }

int lapper_result_copy(const lapper_result_t *r, uint64_t *offsets_out, uint32_t *indices_out) {
    return guarded([&] {
        LAPPER_REQUIRE(r, "null result");
        DeviceGuard dg(r->device);
        // (pageable destinations of some size go through the library's pinned staging buffers: copy_to_host)
        if (offsets_out) copy_to_host(offsets_out, r->d_offsets, (r->nq + 1) * 8);
        if (indices_out && r->total) copy_to_host(indices_out, r->d_indices, r->total * 4);
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
    cudaDeviceSynchronize();  // like cudaFree: readers of the result finish first
    free_result_arrays(r, nullptr);
    if (prev >= 0) cudaSetDevice(prev);
    delete r;
}

int lapper_find_batch_offsets_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                      uint64_t nq, uint64_t *d_offsets, uint64_t *total_out, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(d_offsets && (nq == 0 || (d_q_start && d_q_stop)), "find_batch_offsets: null pointer");
        DeviceGuard dg(l->device);
        ensure_tables(l, (cudaStream_t)stream);
        const uint64_t t = launch_find_offsets(l->ix, d_q_start, d_q_stop, nq, d_offsets, (cudaStream_t)stream);
        if (total_out) *total_out = t;
    });
}

int lapper_find_batch_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                              uint64_t *d_offsets, uint32_t *d_indices, uint64_t indices_capacity, uint64_t *total_out,
                              void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(d_offsets && total_out && (nq == 0 || (d_q_start && d_q_stop)), "find_batch_dev: null pointer");
        LAPPER_REQUIRE(d_indices || indices_capacity == 0, "find_batch_dev: null position buffer with a capacity");
        DeviceGuard dg(l->device);
        ensure_tables(l, (cudaStream_t)stream);
        *total_out = 0;
        const uint64_t total = launch_find(l->ix, d_q_start, d_q_stop, nq, d_offsets, d_indices, indices_capacity, (cudaStream_t)stream);
        *total_out = total;
        if (d_indices && total > indices_capacity)
            throw Error(LAPPER_ERR_CAPACITY, "find_batch_dev: " + std::to_string(total) + " positions do not fit the buffer of " +
                                                 std::to_string(indices_capacity) + " (offsets and *total_out are complete)");
    });
}

int lapper_find_batch_fill_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint64_t nq, const uint64_t *d_offsets, uint32_t *d_indices, void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_offsets), "find_batch_fill: null pointer");
        DeviceGuard dg(l->device);
        ensure_tables(l, (cudaStream_t)stream);
        launch_find_fill(l->ix, d_q_start, d_q_stop, nq, d_offsets, d_indices, (cudaStream_t)stream);
    });
}

int lapper_find_batch_fill_total_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                         uint64_t nq, const uint64_t *d_offsets, uint64_t total, uint32_t *d_indices,
                                         void *stream) {
    return guarded([&] {
        check_handle(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_offsets), "find_batch_fill: null pointer");
        DeviceGuard dg(l->device);
        ensure_tables(l, (cudaStream_t)stream);
        launch_find_fill(l->ix, d_q_start, d_q_stop, nq, d_offsets, d_indices, (cudaStream_t)stream, total);
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
        const DeviceIndex &ix = l->ix;
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
        // the new index is complete before it replaces the old one: a failure leaves the handle as it was
        IndexBuilder b(st.s);
        b.ix.n = n + 1;
        b.ix.starts = pool_alloc<uint32_t>(n + 1, b.ix.bytes, st.s);
        b.ix.stops_by_iv = pool_alloc<uint32_t>(n + 1, b.ix.bytes, st.s);
        b.ix.stops_sorted = pool_alloc<uint32_t>(n + 1, b.ix.bytes, st.s);
        b.ix.perm = pool_alloc<uint32_t>(n + 1, b.ix.bytes, st.s);
        insert_into(b.ix.starts, ix.starts, n, pos[0], start, st.s);
        insert_into(b.ix.stops_by_iv, ix.stops_by_iv, n, pos[0], stop, st.s);
        insert_into(b.ix.stops_sorted, ix.stops_sorted, n, pos[1], stop, st.s);
        insert_into(b.ix.perm, ix.perm, n, pos[0], (uint32_t)id, st.s);
        finish_index_core(b.ix, st.s);  // max_len (src/lib.rs:264-267) falls out of it; the query tables follow on first use
        b.commit(l);
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
        DeviceGuard dg(l->device);
        StreamOwner st;
        IndexBuilder b(st.s);
        // (sets with stop < start: the reference's loop is well defined for them -- comparisons only -- and
        // compute_merged runs it literally, setops.cu / merge_sequential.cuh)
        const bool inverted = l->ix.has_inverted;
        const uint64_t m = compute_merged(l->ix, &b.ix.starts, &b.ix.stops_by_iv, &b.ix.perm, st.s);
        b.ix.n = m;
        b.ix.bytes = 3 * 4 * std::max<uint64_t>(m, 1);
        // runs are disjoint and ascending: their stops are already the sorted `stops` (src/lib.rs:396-406 sorts
        // them again, a no-op) -- unless a run kept an inverted interval's stop: sort then
        b.ix.stops_sorted = pool_alloc<uint32_t>(m, b.ix.bytes, st.s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(b.ix.stops_sorted, b.ix.stops_by_iv, m * 4, cudaMemcpyDeviceToDevice, st.s));
        if (inverted && m > 1) {
            Scratch<uint32_t> tmp(m, st.s);
            launch_sort_u32(b.ix.stops_sorted, tmp.p, m, st.s);
        }
        finish_index_core(b.ix, st.s);  // max_len of the runs (src/lib.rs:408-415); the query tables follow on first use
        b.commit(l);
        l->overlaps_merged = true;  // src/lib.rs:382; the cached cov is left alone like the reference
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper_tune_query_length(lapper_t *l, uint32_t typical_query_len) {
    return guarded([&] {
        check_handle(l);
        DeviceGuard dg(l->device);
        StreamOwner st;
        l->ix.query_len_hint = typical_query_len;
        rebuild_packed_table(l->ix, l->build_flags, st.s);  // (not built yet -- after merge / insert: the lazy build reads the hint)
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
                if (nr) LAPPER_CUDA_CHECK(cudaMemcpyAsync(h[i], d[i], nr * 4, cudaMemcpyDeviceToHost, st.s));
            }
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        } catch (...) {
            for (int i = 0; i < 3; i++) {
                free(h[i]);
                pool_free(d[i], st.s);
            }
            throw;
        }
        for (int i = 0; i < 3; i++) pool_free(d[i], st.s);
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
                // peer copies (NVLink) of the four sorted arrays when the source lives on another GPU; the derived
                // arrays are rebuilt locally
                int rc = lapper_from_sorted_dev(l->ix.starts, l->ix.stops_by_iv, l->ix.stops_sorted, l->ix.perm, l->ix.n,
                                                l->overlaps_merged, devices[i], l->build_flags, st.s, &rep);
                if (rc != LAPPER_OK) throw Error(rc, g_last_error);
                rep->cov_is_some = l->cov_is_some;
                rep->cov = l->cov;
                rep->next_input_id = l->next_input_id;
                if (l->ix.query_len_hint) {
                    rc = lapper_tune_query_length(rep, l->ix.query_len_hint);
                    if (rc != LAPPER_OK) {
                        lapper_free(rep);
                        throw Error(rc, g_last_error);
                    }
                }
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
        // device g owns [g*nq/G, (g+1)*nq/G); every replica runs its own chunked H2D / kernel / D2H pipeline,
        // all of them at once
        for_each_replica(grp, [&](size_t g) {
            const uint64_t lo = nq * g / G, hi = nq * (g + 1) / G;
            count_host<uint64_t>(grp->replicas[g], q_start + lo, q_stop + lo, hi - lo, counts_out + lo);
        });
    });
}

// ---- set operations over a group.  cov: the sorted vector is cut into G contiguous chunks, replica g sums
// the closed form over its chunk (every replica holds the whole prefix max, so run heads and ends at the seams need no
// exchange), the G partial sums are added on the host -- one exchange step.  merge_overlaps: replicas only, every
// replica merges its own copy, all at once (50 M intervals take 0.6 ms on one GPU; cutting that into chunks and
// all-gathering the pieces would cost more than it saves).
int lapper_group_cov(const lapper_group_t *grp, uint32_t *cov_out) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty() && cov_out, "group_cov: null pointer");
        const uint64_t G = grp->replicas.size();
        const lapper_t *first = grp->replicas[0];
        if (first->cov_is_some) {  // src/lib.rs:311-316
            *cov_out = first->cov;
            return;
        }
        need_valid_intervals(first, "cov");
        const uint64_t n = first->ix.n;
        std::vector<uint32_t> part(G, 0);
        for_each_replica(grp, [&](size_t g) {
            const lapper_t *l = grp->replicas[g];
            LAPPER_REQUIRE(l->ix.n == n, "group_cov: the replicas differ (mutate a group only through lapper_group_*)");
            DeviceGuard dg(l->device);
            StreamOwner st;
            part[g] = compute_cov_range(l->ix, n * g / G, n * (g + 1) / G, st.s);
        });
        uint32_t cov = 0;
        for (uint32_t p : part) cov += p;
        *cov_out = cov;
    });
}

int lapper_group_set_cov(lapper_group_t *grp, uint32_t *cov_out) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty(), "group_set_cov: null group");
        for (lapper_t *l : grp->replicas) l->cov_is_some = false;  // src/lib.rs:320-324: always recomputed
        uint32_t cov = 0;
        const int rc = lapper_group_cov(grp, &cov);
        if (rc != LAPPER_OK) throw Error(rc, g_last_error);
        for (lapper_t *l : grp->replicas) l->cov = cov, l->cov_is_some = true;
        if (cov_out) *cov_out = cov;
    });
}

int lapper_group_merge_overlaps(lapper_group_t *grp) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty(), "group_merge_overlaps: null group");
        for_each_replica(grp, [&](size_t g) {
            const int rc = lapper_merge_overlaps(grp->replicas[g]);
            if (rc != LAPPER_OK) throw Error(rc, g_last_error);
        });
    });
}

int lapper_group_find_batch_u32(const lapper_group_t *grp, const uint32_t *q_start, const uint32_t *q_stop,
                                uint64_t nq, uint64_t *offsets_out, uint32_t **indices_out, uint64_t *total_out) {
    return guarded([&] {
        LAPPER_REQUIRE(grp && !grp->replicas.empty(), "null group");
        LAPPER_REQUIRE(offsets_out && indices_out && total_out, "group_find_batch: null pointer");
        LAPPER_REQUIRE(nq == 0 || (q_start && q_stop), "group_find_batch: null pointer");
        *indices_out = nullptr;
        *total_out = 0;
        const uint64_t G = grp->replicas.size();
        std::vector<lapper_result_t *> res(G, nullptr);
        uint32_t *idx = nullptr;
        const auto cleanup = [&] {
            for (auto *r : res) lapper_result_free(r);
        };
        try {
            // phase 1, all replicas at once: queries in, offsets + positions on the device
            for_each_replica(grp, [&](size_t g) {
                const uint64_t lo = nq * g / G, hi = nq * (g + 1) / G;
                int rc = lapper_find_batch_u32(grp->replicas[g], q_start + lo, q_stop + lo, hi - lo, &res[g]);
                if (rc != LAPPER_OK) throw Error(rc, g_last_error);
            });
            // phase 2: each shard's CSR rebased (on its device) by the totals of the shards before it, all copies at once
            std::vector<uint64_t> base(G + 1, 0);
            for (uint64_t g = 0; g < G; g++) base[g + 1] = base[g] + res[g]->total;
            const uint64_t total = base[G];
            idx = (uint32_t *)malloc(std::max<uint64_t>(total, 1) * 4);
            if (!idx) throw std::bad_alloc();
            for_each_replica(grp, [&](size_t g) {
                const uint64_t lo = nq * g / G;
                const lapper_result_t *r = res[g];
                DeviceGuard dg(r->device);
                StreamOwner st;
                const uint64_t cnt = r->nq + (g + 1 == G ? 1 : 0);  // the last shard also writes offsets[nq]
                Scratch<uint64_t> tmp(cnt, st.s);
                launch_rebase_offsets(r->d_offsets, tmp.p, cnt, base[g], st.s);
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(offsets_out + lo, tmp.p, cnt * 8, cudaMemcpyDeviceToHost, st.s));
                if (r->total)
                    LAPPER_CUDA_CHECK(cudaMemcpyAsync(idx + base[g], r->d_indices, r->total * 4, cudaMemcpyDeviceToHost, st.s));
                tmp.release();
                LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
            });
            *indices_out = idx;
            *total_out = total;
        } catch (...) {
            free(idx);
            cleanup();
            throw;
        }
        cleanup();
    });
}

}  // extern "C"
