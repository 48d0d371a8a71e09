// besthit.cuh -- K3: per-pool best-hit selection + stable emission in reference order.
//
// Replaces mWriteBamPool (mBamVector.c:342-347), mWriteBestHitBamPoolByMate,
// mWriteBestHitBamPool, mWriteUniqueBestHitBamPool (msam_filter.c:192-263) and the
// pool segmentation of mFilterFile (msam_filter.c:120-125,170).
//
// Pool == maximal run of byte-identical adjacent QNAMEs: between two consecutive
// mapped records m < i the reference flushes iff some record in (m, i] differs from
// prev_read == QNAME(m), which is exactly "the FB_EQPREV chain from m to i is broken"
// (DESIGN.md gives the two-line proof).  So the run head (FB_EQPREV clear) owns the
// pool and walks it; no sort, no hash, no cross-CTA state.
#pragma once
#include "common.cuh"

namespace msg {

struct BestHitParams {
    uint32_t *fb;             // in/out (FB_KEEP written)
    const int32_t *score;
    uint32_t *segcnt;         // out: kept records emitted at this head (0 for non-heads)
    uint64_t n;
    int uniq;
    uint32_t *err;
};

__device__ __forceinline__ int mate_class(uint32_t fb)
{   // 0: neither/both flags... class index by (flag & 0xC0): 0x00->0, 0x40->1, 0x80->2, 0xC0->3
    return (int)((fb >> 6) & 3u);
}

__global__ void __launch_bounds__(256) besthit_select_kernel(const BestHitParams p)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t f0 = p.fb[i];
    if (i > 0 && (f0 & FB_EQPREV)) { p.segcnt[i] = 0; return; }          // not a pool head

    // pass 1: per mate class best score / tie count (msam_filter.c:212-230), pool pairedness (:196-204)
    int32_t best[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};
    int32_t cnt[4] = {0, 0, 0, 0};
    bool noas[4] = {false, false, false, false};
    bool paired = false;
    uint64_t j = i;
    uint32_t f = f0;
    for (;;) {
        if (f & FB_INPOOL) {
            int c = mate_class(f);
            paired |= (c != 0);
            if (!(f & FB_HAS_AS)) noas[c] = true;
            int32_t s = p.score[j];
#pragma unroll
            for (int k = 0; k < 4; k++) if (k == c) {
                if (s > best[k]) { best[k] = s; cnt[k] = 1; } else if (s == best[k]) cnt[k]++;
            }
        }
        if (++j >= p.n) break;
        f = p.fb[j];
        if (!(f & FB_EQPREV)) break;
    }
    const uint64_t end = j;
    // classes the reference visits: READ1 then READ2 when paired, else class 0 (:247-254)
    if (paired ? (noas[1] || noas[2]) : noas[0]) atomicOr(p.err, DERR_NOAS);               // :219-221

    // pass 2: keep bits (:235-244)
    uint32_t kept = 0;
    for (j = i; j < end; j++) {
        f = p.fb[j];
        bool keep = false;
        if (f & FB_INPOOL) {
            int c = mate_class(f);
            bool act = paired ? (c == 1 || c == 2) : (c == 0);
            int32_t s = p.score[j];
            int32_t b = INT32_MIN, n = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if (k == c) { b = best[k]; n = cnt[k]; }
            keep = act && s == b && (!p.uniq || n == 1);
        }
        uint32_t nf = keep ? (f | FB_KEEP) : (f & ~FB_KEEP);
        if (nf != f) p.fb[j] = nf;
        kept += keep;
    }
    p.segcnt[i] = kept;
}

// heads write their pool's winners: all READ1-class (or unpaired) winners in input
// order, then all READ2-class winners (msam_filter.c:247-254).
__global__ void __launch_bounds__(256) besthit_emit_kernel(const uint32_t *fb, const uint32_t *segcnt, const uint32_t *segbase,
                                                           uint32_t *out_idx, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (segcnt[i] == 0) return;
    uint32_t c1 = 0;
    uint64_t j = i;
    for (;;) {
        uint32_t f = fb[j];
        if (j > i && !(f & FB_EQPREV)) break;
        if ((f & FB_KEEP) && mate_class(f) != 2) c1++;
        if (++j >= n) break;
    }
    const uint64_t end = j;
    uint32_t w1 = segbase[i], w2 = w1 + c1;
    for (j = i; j < end; j++) {
        uint32_t f = fb[j];
        if (f & FB_KEEP) { if (mate_class(f) == 2) out_idx[w2++] = (uint32_t)j; else out_idx[w1++] = (uint32_t)j; }
    }
}

// ---- scan functors -----------------------------------------------------------------------
struct InFlagBit {           // 1 where (fb & mask) == want
    const uint32_t *fb; uint32_t mask, want;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return (fb[i] & mask) == want ? 1u : 0u; }
};
struct InU32 { const uint32_t *v; __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return v[i]; } };
struct OutExclU32 { uint32_t *o; __device__ __forceinline__ void operator()(uint64_t i, uint32_t ex, uint32_t) const { o[i] = ex; } };
struct OutInclU32 { uint32_t *o; __device__ __forceinline__ void operator()(uint64_t i, uint32_t ex, uint32_t v) const { o[i] = ex + v; } };
struct OutCompact {          // stable compaction: out[ex] = i where the flag was set
    uint32_t *o; __device__ __forceinline__ void operator()(uint64_t i, uint32_t ex, uint32_t v) const { if (v) o[ex] = (uint32_t)i; }
};

} // namespace msg
