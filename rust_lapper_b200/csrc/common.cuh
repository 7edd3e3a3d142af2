// common.cuh -- shared host/device helpers for liblapper_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/lapper_b200.h"

namespace lap {

// ---------------------------------------------------------------- errors (never cross the C ABI)
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define LAPPER_CUDA_CHECK(expr)                                                                            \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            int _code = (_e == cudaErrorMemoryAllocation) ? LAPPER_ERR_OOM : LAPPER_ERR_CUDA;              \
            throw ::lap::Error(_code, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                                             __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
        }                                                                                                  \
    } while (0)

#define LAPPER_REQUIRE(cond, msg)                                                     \
    do {                                                                              \
        if (!(cond)) throw ::lap::Error(LAPPER_ERR_INVALID_ARG, std::string(msg)); \
    } while (0)

__host__ __device__ static inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// Device selection guard (restores the caller's current device).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        LAPPER_CUDA_CHECK(cudaGetDevice(&prev));
        if (prev != dev) LAPPER_CUDA_CHECK(cudaSetDevice(dev));
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Stream-ordered scratch memory from a library-owned pool (one per device).  The release threshold is
// "never": a batch call that needs 400 MB of scratch gets it back from the pool on the next call
// instead of from the driver (which costs about a millisecond per call).
cudaMemPool_t scratch_pool();  // api.cu
// number of SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device): persistent grids are sized
// as a multiple of it (148 on B200)
uint32_t sm_count();  // api.cu

template <typename T>
struct Scratch {
    T *p = nullptr;
    cudaStream_t s = nullptr;
    Scratch() = default;
    Scratch(uint64_t count, cudaStream_t stream) { alloc(count, stream); }
    void alloc(uint64_t count, cudaStream_t stream) {
        release();
        s = stream;
        if (count == 0) count = 1;
        LAPPER_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&p, count * sizeof(T), scratch_pool(), stream));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    T *take() {
        T *q = p;
        p = nullptr;
        return q;
    }
    ~Scratch() { release(); }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// compute-sanitizer is not available on the GPU pool, so the kernels carry their own bounds checks: a build with
// -DLAPPER_DEBUG_CHECKS counts violated conditions in a device counter (lapper_debug_violations(), api.cu) instead of
// trusting the output comparison alone; the GPU test suite is run against that build (tests/conftest.py reports it).
#ifdef LAPPER_DEBUG_CHECKS
// (the counter lives in query.cu, the only file whose kernels carry checks: no relocatable device code needed)
#define LAPPER_CHECK(cond)                                   \
    do {                                                     \
        if (!(cond)) atomicAdd(&::lap::g_debug_violations, 1ull); \
    } while (0)
#else
#define LAPPER_CHECK(cond) \
    do {                   \
    } while (0)
#endif

struct __align__(32) Sector {
    uint32_t w[8];
};

// One 32-byte sector per instruction (LDG.E.256 on sm_100a), read-only path, no L1 allocation:
// a random bucket is touched once per query, L1 has nothing to add for it.
__device__ __forceinline__ Sector ld_sector_stream(const void *p) {
    Sector r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
    return r;
}
// Same sector, allocating in L1 (start-sorted query batches revisit neighbouring buckets).
__device__ __forceinline__ Sector ld_sector_cached(const void *p) {
    Sector r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
    return r;
}

// Streaming query loads / result stores: read or written exactly once, keep them out of the way of
// the index in L1/L2 (.cs = cache-streaming, evict-first).
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u64(void *p, uint64_t v) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}



// ---- L2 residency control (createpolicy + per-access cache hints)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ Sector ld_sector_hint(const void *p, uint64_t pol) {
    Sector r;
    asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ld_v4_hint(const void *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_v4_hint(void *p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}

// 256-bit streaming accesses: eight u32 per lane, read or written once (no L1 allocation)
__device__ __forceinline__ Sector ld_stream_v8(const void *p) { return ld_sector_stream(p); }
__device__ __forceinline__ void st_stream_v8(void *p, const Sector &v) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]),
                 "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7])
                 : "memory");
}

// lower_bound over a sorted u32 range [lo, hi): first index whose element is >= key.
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    uint32_t key) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
// upper_bound: first index whose element is > key.
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    uint32_t key) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

#endif  // __CUDACC__

}  // namespace lap
