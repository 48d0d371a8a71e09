// coverage.cuh -- K6: per-position depth as a difference array, then per-target summary.
//
// Replaces mUpdateCoverageForAlignment (msam_coverage.c:33-87; the reference's inner
// loop is per *base*, ours is two atomics per M/=/X run) and the numeric part of
// mWriteCoverageSummaryToStream (msam_coverage.c:189-219).
//
// Layout: one int32 array `diff` of sum(tlen[t] + 1) cells; target t owns
// [covbase[t], covbase[t] + tlen[t]] -- the extra cell absorbs the "-1" of runs that
// end exactly at tlen, so every target's cells sum to zero and ONE unsegmented
// prefix sum over the whole array yields the depth of every target.
#pragma once
#include "common.cuh"
#include "decode.cuh"

namespace msg {

// stream-driven coverage (used when the kept list is only known after best-hit selection)
__global__ void __launch_bounds__(256) coverage_stream_kernel(const uint8_t *raw, const uint64_t *off, const uint32_t *stream, uint64_t m,
                                                              int32_t n_targets, int32_t *diff, const uint64_t *covbase,
                                                              const uint32_t *tlen, uint8_t *covered, uint32_t *err,
                                                              unsigned long long *covbits, unsigned long long *covsum)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t r = stream ? stream[j] : (uint32_t)j;
    const uint64_t o = off[r];
    GlAcc g{raw + o};
    RecCore c = parse_core(g, off[r + 1] - o);
    if (c.bad) { atomicOr(err, DERR_FORMAT); return; }
    if (c.tid < 0) return;                                              // msam_coverage.c:42
    if (c.tid >= n_targets) { atomicOr(err, DERR_FORMAT); return; }
    if (covbits) cover_record_bits(g, 36 + c.lq, c.nc, c.tid, c.pos, covbits, covbase, tlen, covered, covsum);
    else cover_record(g, 36 + c.lq, c.nc, c.tid, c.pos, diff, covbase, tlen, covered);
}

// after the prefix sum: per-target (#cells != 0, sum of cells) over [covbase[t], covbase[t]+tlen[t])
// Each thread owns SPAN consecutive cells; target lookup by binary search on covbase.
constexpr int COV_SPAN = 16;

__device__ __forceinline__ void cov_flush(unsigned long long *touched, long long *sum, int32_t t, unsigned long long tc, long long sm)
{
    // consecutive threads cover consecutive cells, so a warp usually sits inside ONE target: combine the 32 partial
    // results first and issue one pair of atomics per warp instead of per thread (a few hundred targets would otherwise
    // serialise tens of millions of same-address atomics).  Called by all 32 lanes (t = -1: nothing to add).
    const int32_t t0 = __shfl_sync(0xffffffffu, t, 0);
    if (__all_sync(0xffffffffu, t == t0)) {
        if (t0 < 0) return;
        tc = warp_sum_u64(tc);
        sm = (long long)warp_sum_u64((unsigned long long)sm);
        if ((threadIdx.x & 31u) == 0 && (tc | (unsigned long long)sm)) {
            atomicAdd(touched + t, tc); if (sum) atomicAdd((unsigned long long *)(sum + t), (unsigned long long)sm);
        }
    } else if (t >= 0 && (tc | (unsigned long long)sm)) {
        atomicAdd(touched + t, tc); if (sum) atomicAdd((unsigned long long *)(sum + t), (unsigned long long)sm);
    }
}

__global__ void __launch_bounds__(256) coverage_reduce_kernel(const int32_t *depth, uint64_t total_cells, const uint64_t *covbase,
                                                              const uint32_t *tlen, int32_t n_targets,
                                                              unsigned long long *touched, long long *sum)
{
    const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * COV_SPAN;
    int32_t t = -1;
    unsigned long long tc = 0; long long sm = 0;
    if (start < total_cells) {
        const uint64_t stop = min(total_cells, start + COV_SPAN);
        // largest t with covbase[t] <= start
        int32_t lo = 0, hi = n_targets - 1;
        while (lo < hi) { int32_t mid = (lo + hi + 1) >> 1; if (covbase[mid] <= start) lo = mid; else hi = mid - 1; }
        t = lo;
        uint64_t tend = covbase[t] + tlen[t];          // first cell that is NOT a position of t (the spill cell)
        for (uint64_t x = start; x < stop; x++) {
            while (x > tend) {                         // moved past t's spill cell: rare (target boundary inside my span)
                if (tc | (unsigned long long)sm) { atomicAdd(touched + t, tc); atomicAdd((unsigned long long *)(sum + t), (unsigned long long)sm); }
                tc = 0; sm = 0; t++; tend = covbase[t] + tlen[t];
            }
            if (x == tend) continue;                   // spill cell
            int32_t v = depth[x];
            tc += (v != 0); sm += v;
        }
    }
    cov_flush(touched, sum, t, tc, sm);
}

// summary mode: touched[t] = popcount of target t's bitmap.  Each thread owns 4 consecutive 64-bit words; targets start
// on word boundaries (wbase), bits past tlen are never set.
__global__ void __launch_bounds__(256) coverage_popcount_kernel(const unsigned long long *bits, uint64_t total_words, const uint64_t *wbase,
                                                                int32_t n_targets, unsigned long long *touched)
{
    const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int32_t t = -1;
    unsigned long long tc = 0;
    if (start < total_words) {
        const uint64_t stop = min(total_words, start + 4);
        int32_t lo = 0, hi = n_targets - 1;
        while (lo < hi) { int32_t mid = (lo + hi + 1) >> 1; if (wbase[mid] <= start) lo = mid; else hi = mid - 1; }
        t = lo;
        uint64_t tend = wbase[t + 1];
        for (uint64_t x = start; x < stop; x++) {
            while (x >= tend) {                        // crossed into the next target (possibly over empty ones)
                if (tc) atomicAdd(touched + t, tc);
                tc = 0; t++; tend = wbase[t + 1];
            }
            tc += (unsigned long long)__popcll(bits[x]);
        }
    }
    cov_flush(touched, nullptr, t, tc, 0);
}

struct InI32 { const int32_t *v; __device__ __forceinline__ int32_t operator()(uint64_t i) const { return v[i]; } };
struct OutInclI32 { int32_t *o; __device__ __forceinline__ void operator()(uint64_t i, int32_t ex, int32_t v) const { o[i] = ex + v; } };

} // namespace msg
