// Microbenchmark: random 32-byte sector gathers (LDG.E.256) from a table of a given size, as a function
// of table size, independent loads in flight per thread and occupancy.  Sets the hardware ceiling for the
// bucket-sector count kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_gather mb_gather.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

struct __align__(32) Sector { uint32_t w[8]; };

__device__ __forceinline__ Sector ld256(const void* p) {
    Sector r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}

// each thread performs `iters` rounds of K independent gathers; also streams `stream_bytes_per_gather` optional
// g_split: 0 = every CTA gathers from the whole table; otherwise a CTA gathers from ONE half of it only, chosen by
// 1 = parity of %smid, 2 = %smid below / above half the SM count, 3 = parity of blockIdx.x (control: no relation to
// where the CTA runs), 4 = (%smid / 2) parity (TPC pairs).  Tests whether a table half that is only ever read by the SMs
// of one die doubles the L2 capacity that is effective for the table.
__constant__ int g_split;
template <int K, bool STREAM>
__global__ void gather_kernel(const Sector* __restrict__ table, uint32_t nsect, int iters, uint32_t* __restrict__ out,
                              const uint4* __restrict__ stream_in, uint32_t* __restrict__ stream_out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    if (g_split) {
        uint32_t smid, nsm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        const uint32_t half = g_split == 1 ? (smid & 1u) : g_split == 2 ? (smid >= nsm / 2 ? 1u : 0u) : g_split == 3 ? (blockIdx.x & 1u) : ((smid >> 1) & 1u);
        nsect /= 2;
        table += (size_t)half * nsect;
    }
    for (int it = 0; it < iters; it++) {
        const uint64_t q0 = ((uint64_t)it * nthreads + tid) * K;
        uint32_t extra = 0;
        if (STREAM) {  // 8 B in per gather (coalesced), like the query arrays
            const uint4* p = stream_in + q0 * 8 / 16;
#pragma unroll
            for (int k = 0; k < K / 2; k++) { uint4 v = __ldcs(p + k); extra += v.x ^ v.y ^ v.z ^ v.w; }
        }
        Sector s[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            uint32_t idx = (uint32_t)(((uint64_t)mix(q0 + k) * nsect) >> 32);
            s[k] = ld256(table + idx);
        }
#pragma unroll
        for (int k = 0; k < K; k++) acc += s[k].w[0] + s[k].w[7];
        acc += extra;
        if (STREAM) {
#pragma unroll
            for (int k = 0; k < K; k++) __stcs(stream_out + q0 + k, acc + k);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// The same work with the streams moved by the TMA path instead of LSU instructions: a CTA's tile of 1024 "queries"
// (8 B each) arrives in shared memory by one bulk copy, the results (4 B each) leave by one bulk store; only the
// gathers go through the L1TEX tag stage.  Tests whether the sector rate of that stage (one per clock per SM) is what
// the streams' LDG / STG instructions compete for.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int K>
__global__ void __launch_bounds__(256) gather_tma_kernel(const Sector* __restrict__ table, uint32_t nsect, int iters,
                                                         uint32_t* __restrict__ out, const uint4* __restrict__ stream_in,
                                                         uint32_t* __restrict__ stream_out) {
    // per-warp pipelines (no CTA barrier): a warp's 32 * K queries = one bulk load of 32 * K * 8 bytes, two stages;
    // its 32 * K results = one bulk store, two buffers.  Streams carry the L2 evict-first policy like the real kernel's.
    struct __align__(128) WarpBuf {
        uint32_t qin[2][32 * K * 2];
        uint32_t res[2][32 * K];
        uint64_t full[2];
        uint64_t pad[14];
    };
    __shared__ WarpBuf bufs[8];
    WarpBuf& wb = bufs[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_global = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    uint32_t acc = 0;
    if (g_split) {
        uint32_t smid, nsm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        const uint32_t half = g_split == 1 ? (smid & 1u) : (smid >= nsm / 2 ? 1u : 0u);
        nsect /= 2;
        table += (size_t)half * nsect;
    }
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (lane == 0) {
        for (int s = 0; s < 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&wb.full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t tile_bytes = 32 * K * 8;
    auto fetch = [&](int it) {
        const uint64_t q0 = ((uint64_t)it * nwarps + warp_global) * (32 * K);
        const int s = it & 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&wb.full[s])), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         smem_addr(wb.qin[s])),
                     "l"((const char*)stream_in + q0 * 8), "r"(tile_bytes), "r"(smem_addr(&wb.full[s])), "l"(pol)
                     : "memory");
    };
    if (lane == 0) fetch(0);
    for (int it = 0; it < iters; it++) {
        const int s = it & 1;
        if (lane == 0 && it + 1 < iters) fetch(it + 1);
        {
            uint32_t ok;
            const uint32_t parity = (it >> 1) & 1;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_addr(&wb.full[s])), "r"(parity) : "memory");
            } while (!ok);
        }
        const uint64_t q0 = ((uint64_t)it * nwarps + warp_global) * (32 * K) + lane * K;
        uint32_t extra = 0;
#pragma unroll
        for (int k = 0; k < K / 2; k++) {
            const uint4 v = *reinterpret_cast<const uint4*>(&wb.qin[s][(lane * K + 2 * k) * 2]);
            extra += v.x ^ v.y ^ v.z ^ v.w;
        }
        Sector sc[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            uint32_t idx = (uint32_t)(((uint64_t)mix(q0 + k) * nsect) >> 32);
            sc[k] = ld256(table + idx);
        }
#pragma unroll
        for (int k = 0; k < K; k++) acc += sc[k].w[0] + sc[k].w[7];
        acc += extra;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        uint4 r;
        r.x = acc, r.y = acc + 1, r.z = acc + 2, r.w = acc + 3;
        static_assert(K == 4, "one 128-bit result store per thread");
        *reinterpret_cast<uint4*>(&wb.res[s][lane * K]) = r;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint64_t o0 = ((uint64_t)it * nwarps + warp_global) * (32 * K);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(stream_out + o0),
                         "r"(smem_addr(wb.res[s])), "r"(32 * K * 4), "l"(pol)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) out[0] = acc;
}

float run_tma(const Sector* table, uint32_t nsect, uint64_t total_gathers, int blocks, uint32_t* out, const uint4* sin, uint32_t* sout) {
    const int K = 4, threads = 256;
    int iters = (int)(total_gathers / ((uint64_t)threads * blocks * K));
    if (iters < 1) iters = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather_tma_kernel<K><<<blocks, threads>>>(table, nsect, iters, out, sin, sout);
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) gather_tma_kernel<K><<<blocks, threads>>>(table, nsect, iters, out, sin, sout);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    double g = (double)iters * threads * blocks * K;
    printf("K=%d stream=TMA threads=%d blocks=%d : %.3f ms  %.1f Ggather/s  %.1f GB/s(sector bytes)  [%s]\n", K, threads, blocks, ms,
           g / ms * 1e-6, g * 32 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    return ms;
}

template <int K, bool STREAM>
float run(const Sector* table, uint32_t nsect, uint64_t total_gathers, int threads, int blocks, uint32_t* out,
          const uint4* sin, uint32_t* sout) {
    int iters = (int)(total_gathers / ((uint64_t)threads * blocks * K));
    if (iters < 1) iters = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<K, STREAM><<<blocks, threads>>>(table, nsect, iters, out, sin, sout);
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) gather_kernel<K, STREAM><<<blocks, threads>>>(table, nsect, iters, out, sin, sout);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    double g = (double)iters * threads * blocks * K;
    printf("K=%d stream=%d threads=%d blocks=%d : %.3f ms  %.1f Ggather/s  %.1f GB/s(sector bytes)\n", K, (int)STREAM, threads, blocks, ms,
           g / ms * 1e-6, g * 32 / ms * 1e-6);
    return ms;
}

int main(int argc, char** argv) {
    const uint64_t total = 100000000ull;
    uint32_t* out; cudaMalloc(&out, 4);
    uint4* sin; uint32_t* sout; cudaMalloc(&sin, total * 8 + 4096); cudaMalloc(&sout, total * 4 + 4096);
    cudaMemset(sin, 1, total * 8);
    if (argc > 1 && !strcmp(argv[1], "one")) {
        // mb_gather one <MB> <split>: a single configuration (4 gathers in flight per thread, query / result streams,
        // 4 CTAs of 256 threads per SM) -- the one to put under ncu
        const double mb = argc > 2 ? atof(argv[2]) : 75;
        const int split = argc > 3 ? atoi(argv[3]) : 0;
        cudaMemcpyToSymbol(g_split, &split, sizeof(int));
        uint32_t nsect = (uint32_t)(mb * 1e6 / 32);
        Sector* table; cudaMalloc(&table, (size_t)nsect * 32); cudaMemset(table, 1, (size_t)nsect * 32);
        int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
        printf("== table %.0f MB (%u sectors), split mode %d\n", mb, nsect, split);
        run<4, true>(table, nsect, total, 256, nsm * 4, out, sin, sout);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "tma")) {
        // streams by LSU instructions vs by bulk copies (TMA), table sizes on both sides of the L2 knee
        int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
        { const int zero = 0; cudaMemcpyToSymbol(g_split, &zero, sizeof(int)); }
        double sizes[] = {12, 48, 64, 75, 97};
        for (double mb : sizes) {
            uint32_t nsect = (uint32_t)(mb * 1e6 / 32);
            Sector* table; cudaMalloc(&table, (size_t)nsect * 32); cudaMemset(table, 1, (size_t)nsect * 32);
            for (int occ : {3, 4, 6, 8}) {
                printf("== table %.0f MB, %d CTAs/SM: ", mb, occ);
                run<4, true>(table, nsect, total, 256, nsm * occ, out, sin, sout);
                printf("== table %.0f MB, %d CTAs/SM: ", mb, occ);
                run_tma(table, nsect, total, nsm * occ, out, sin, sout);
            }
            cudaFree(table);
        }
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "split")) {
        // the die-split experiment: table sizes around and beyond the L2 knee, every split mode, with and without streams
        int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
        double sizes[] = {48, 75, 97, 128, 150, 194};
        for (double mb : sizes) {
            uint32_t nsect = (uint32_t)(mb * 1e6 / 32);
            Sector* table; cudaMalloc(&table, (size_t)nsect * 32); cudaMemset(table, 1, (size_t)nsect * 32);
            for (int split = 0; split <= 4; split++) {
                cudaMemcpyToSymbol(g_split, &split, sizeof(int));
                printf("== table %.0f MB, split mode %d: ", mb, split);
                run<4, false>(table, nsect, total, 256, nsm * 4, out, sin, sout);
                printf("== table %.0f MB, split mode %d: ", mb, split);
                run<4, true>(table, nsect, total, 256, nsm * 4, out, sin, sout);
            }
            cudaFree(table);
        }
        return 0;
    }
    { const int zero = 0; cudaMemcpyToSymbol(g_split, &zero, sizeof(int)); }
    double sizes_mb[] = {12, 24, 48, 64, 80, 97, 194, 485};
    for (double mb : sizes_mb) {
        uint32_t nsect = (uint32_t)(mb * 1e6 / 32);
        Sector* table; cudaMalloc(&table, (size_t)nsect * 32); cudaMemset(table, 1, (size_t)nsect * 32);
        printf("== table %.0f MB (%u sectors)\n", mb, nsect);
        for (int occ : {2, 4, 8}) {
            int blocks = 148 * occ;
            run<1, false>(table, nsect, total, 256, blocks, out, sin, sout);
            run<4, false>(table, nsect, total, 256, blocks, out, sin, sout);
            run<8, false>(table, nsect, total, 256, blocks, out, sin, sout);
            run<4, true>(table, nsect, total, 256, blocks, out, sin, sout);
            run<8, true>(table, nsect, total, 256, blocks, out, sin, sout);
        }
        cudaFree(table);
    }
    return 0;
}
