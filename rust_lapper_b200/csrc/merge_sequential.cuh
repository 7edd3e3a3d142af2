// merge_sequential.cuh -- Lapper::merge_overlaps (src/lib.rs:361-416) as the reference's own loop, for sets in which
// some interval has stop < start.
//
// For start <= stop sets the running `top.stop` of the reference's stack is the prefix max of the stops, and the merge
// is a scan (setops.cu, engine64.cu).  With an inverted interval at the head of a run that identity fails -- the run's
// `top.stop` restarts BELOW stops seen earlier -- and the state transition  top.stop < start ? (new run) : max(top.stop,
// stop)  does not compose into a bounded-size scan operator.  Such sets are degenerate input; they get the literal
// loop, run by one warp: 32 intervals per coalesced load, the lanes step through them together (every lane holds the
// same state), lane 0 writes.  About 10 ns per interval -- slow next to the scan, exact like the reference.
#pragma once
#include <cstdint>

namespace lap {

// kWrite = false: counts the runs (*count_out); kWrite = true: writes (start, stop, perm of the run's first interval)
template <typename I, bool kWrite>
__global__ void __launch_bounds__(32) merge_sequential_kernel(const I *__restrict__ S, const I *__restrict__ E,
                                                              const uint32_t *__restrict__ perm, uint64_t n,
                                                              I *__restrict__ out_s, I *__restrict__ out_e,
                                                              uint32_t *__restrict__ out_perm,
                                                              unsigned long long *__restrict__ count_out) {
    const uint32_t lane = threadIdx.x;
    uint64_t m = 0;
    I top_s = 0, top_e = 0;
    uint32_t top_p = 0;
    bool have = false;
    for (uint64_t base = 0; base < n; base += 32) {
        const uint64_t i = base + lane;
        I s = 0, e = 0;
        uint32_t p = 0;
        if (i < n) s = S[i], e = E[i], p = perm[i];
        const int cnt = (int)(n - base < 32 ? n - base : 32);
        for (int k = 0; k < cnt; k++) {
            const I sk = __shfl_sync(0xFFFFFFFFu, s, k), ek = __shfl_sync(0xFFFFFFFFu, e, k);
            const uint32_t pk = __shfl_sync(0xFFFFFFFFu, p, k);
            if (!have) {  // :366 the first interval opens the stack
                top_s = sk, top_e = ek, top_p = pk, have = true;
            } else if (top_e < sk) {  // :372-374 both stay: the top is final
                if (kWrite && lane == 0) out_s[m] = top_s, out_e[m] = top_e, out_perm[m] = top_p;
                m++;
                top_s = sk, top_e = ek, top_p = pk;
            } else if (top_e < ek) {  // :375-377 the top grows
                top_e = ek;
            }  // :378-380 swallowed
        }
    }
    if (have) {
        if (kWrite && lane == 0) out_s[m] = top_s, out_e[m] = top_e, out_perm[m] = top_p;
        m++;
    }
    if (!kWrite && lane == 0) *count_out = m;
}

}  // namespace lap
