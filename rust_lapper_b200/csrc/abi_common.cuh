// abi_common.cuh -- what the files that define C-ABI entry points share (api.cu: I = u32, engine64.cu: I = u64): the
// exception -> status code boundary, the per-thread error text behind lapper_last_error(), device checks.
#pragma once

#include <cstdlib>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace lap {

std::string &last_error_slot();  // api.cu: the calling thread's last error text
// api.cu: copies between device memory and (possibly pageable) host memory; large pageable transfers are staged through
// the library's pinned buffers by copy threads.  copy_to_host is synchronous; copy_to_device is stream-ordered for
// pinned / small sources and returns after the data has left the caller's memory either way.
void copy_to_host(void *h_dst, const void *d_src, size_t bytes);
void copy_to_device(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream);
// the pieces of that staging, for the host pipelines that overlap it with kernels (count_host_staged and its u64 twin)
bool host_pointer_is_pinned(const void *p);
int host_copy_threads();  // LAPPER_HOST_COPY_THREADS (default 4; 0 = leave pageable buffers to the driver)
void parallel_memcpy(void *dst, const void *src, size_t bytes);
struct PinnedBuf {  // a pinned staging buffer from the process-wide cache (returned to it on destruction)
    void *p = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    explicit PinnedBuf(size_t n);
    ~PinnedBuf();
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    PinnedBuf(PinnedBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; }
};
void require_device_index(int device);  // api.cu: throws LAPPER_ERR_NO_DEVICE / LAPPER_ERR_INVALID_ARG

template <typename F>
int guarded_call(F &&f) {
    try {
        f();
        return LAPPER_OK;
    } catch (const Error &e) {
        last_error_slot() = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        last_error_slot() = "host allocation failed";
        return LAPPER_ERR_OOM;
    } catch (const std::exception &e) {
        last_error_slot() = e.what();
        return LAPPER_ERR_CUDA;
    } catch (...) {
        last_error_slot() = "unknown error";
        return LAPPER_ERR_CUDA;
    }
}

// A non-blocking stream for the duration of one call.  Creating and destroying one costs about 100 us here -- several
// times the rest of a single-query call -- so the process keeps the streams its calls have used (per device) and hands
// them out again: a call's stream is still its own (nothing else runs on it until the call returns it), and work left
// on it by a call that threw is simply ordered before the next call's.  One list for all host threads (the replica
// groups run every call on fresh threads); never destroyed (the driver may be gone by static destruction time).
struct StreamCache {
    std::mutex mu;
    std::vector<std::pair<int, cudaStream_t>> free_list;
    cudaStream_t get(int device) {
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = free_list.size(); i-- > 0;)
            if (free_list[i].first == device) {
                cudaStream_t s = free_list[i].second;
                free_list.erase(free_list.begin() + (long)i);
                return s;
            }
        return nullptr;
    }
    void put(int device, cudaStream_t s) {
        std::lock_guard<std::mutex> g(mu);
        free_list.emplace_back(device, s);
    }
};
inline StreamCache &stream_cache() {
    static StreamCache *c = new StreamCache();
    return *c;
}
inline bool stream_cache_enabled() {  // LAPPER_NO_STREAM_CACHE=1: a fresh stream per call, destroyed at its end
    static const bool on = [] {
        const char *e = std::getenv("LAPPER_NO_STREAM_CACHE");
        return !(e && *e && *e != '0');
    }();
    return on;
}
struct OwnedStream {
    cudaStream_t s = nullptr;
    int device = 0;
    OwnedStream() {
        LAPPER_CUDA_CHECK(cudaGetDevice(&device));
        if (stream_cache_enabled()) s = stream_cache().get(device);
        if (!s) LAPPER_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
    ~OwnedStream() {
        if (s && stream_cache_enabled())
            stream_cache().put(device, s);
        else if (s)
            cudaStreamDestroy(s);
    }
    OwnedStream(const OwnedStream &) = delete;
    OwnedStream &operator=(const OwnedStream &) = delete;
};

}  // namespace lap
