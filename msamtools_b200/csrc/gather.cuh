// gather.cuh -- materialise the filter stage's output records (mSamWrite of every
// kept record, msam_filter.c:243 / mBamVector.c:342-347) in reference output order.
// With --rescore the first AS field is removed and "AS:i:<score>" appended, exactly
// what bam_aux_del + bam_aux_append do at msam_filter.c:160-168.
#pragma once
#include "common.cuh"
#include "decode.cuh"

namespace msg {

struct GatherPlan {          // per stream record
    unsigned long long src;  // byte offset of the record in the chunk (so that the copy kernel needs no stream[] -> off[] chain)
    uint32_t len;            // bytes of the source record
    uint32_t cut0, cut1;     // byte span [cut0,cut1) of the record to drop (AS field), cut0==cut1 -> none
    uint32_t append;         // 1 -> append AS:i
};

// thread per stream record: output length + rescore plan
__global__ void __launch_bounds__(256) gather_plan_kernel(const uint8_t *raw, const uint64_t *off, const uint32_t *stream, uint64_t m,
                                                          int rescore, uint32_t *out_len, GatherPlan *plan)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t r = stream ? stream[j] : (uint32_t)j;
    const uint64_t o = off[r];
    const uint32_t len = (uint32_t)(off[r + 1] - o);
    uint32_t nl = len;
    GatherPlan pl = {o, len, 0, 0, 0};
    if (rescore) {
        GlAcc g{raw + o};
        RecCore c = parse_core(g, len);
        if (!c.bad && !(c.flag & BAM_FUNMAP)) {              // unmapped records are never rescored (:132-138 precede :160)
            // locate the first AS field (bam_aux_get), same walk as aux_scan
            GlAcc ax{raw + o + c.aux_off};
            uint32_t y = 0;
            while (y + 3 <= c.aux_len) {
                uint32_t t0 = ax.u8(y), t1 = ax.u8(y + 1), ty = ax.u8(y + 2);
                uint32_t v0 = y + 3, vend = 0;
                if (ty == 'A' || ty == 'c' || ty == 'C') vend = v0 + 1;
                else if (ty == 's' || ty == 'S') vend = v0 + 2;
                else if (ty == 'i' || ty == 'I' || ty == 'f') vend = v0 + 4;
                else if (ty == 'd') vend = v0 + 8;
                else if (ty == 'Z' || ty == 'H') {
                    uint32_t z = v0; while (z < c.aux_len && ax.u8(z)) z++;
                    if (z >= c.aux_len) break;
                    vend = z + 1;
                } else if (ty == 'B') {
                    if (v0 + 5 > c.aux_len) break;
                    uint32_t st = ax.u8(v0), cnt = ax.u32(v0 + 1), es;
                    if (st == 'c' || st == 'C') es = 1; else if (st == 's' || st == 'S') es = 2;
                    else if (st == 'i' || st == 'I' || st == 'f') es = 4; else break;
                    unsigned long long tot = (unsigned long long)v0 + 5ull + (unsigned long long)cnt * es;
                    if (tot > c.aux_len) break;
                    vend = (uint32_t)tot;
                } else break;
                if (vend > c.aux_len) break;
                if (t0 == 'A' && t1 == 'S') { pl.cut0 = c.aux_off + y; pl.cut1 = c.aux_off + vend; break; }
                y = vend;
            }
            pl.append = 1;
            nl = len - (pl.cut1 - pl.cut0) + 7;
        }
    }
    out_len[j] = nl;
    plan[j] = pl;
}

// warp per stream record.  Source and destination are both byte-aligned at best, with different misalignments: the copy
// runs on the DESTINATION's 32-bit grid -- every lane stores one aligned word assembled from two aligned source words
// with a funnel shift (neighbouring lanes share those loads in L1), so a 300-byte record is three coalesced 128-byte
// store rounds instead of ten rounds of single bytes; only the <= 3 bytes before the first and after the last whole
// destination word are copied bytewise.  --rescore records (AS field cut out, AS:i appended, block_size patched) keep
// the byte loop.
__global__ void __launch_bounds__(256) gather_copy_kernel(const uint8_t *raw, const uint64_t *off, const uint32_t *stream, uint64_t m,
                                                          const unsigned long long *out_off, const uint32_t *out_len,
                                                          const GatherPlan *plan, const int32_t *score, uint8_t *out)
{
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (j >= m) return;
    const GatherPlan pl = plan[j];
    const uint8_t *src = raw + pl.src;
    const uint32_t len = pl.len;
    uint8_t *dst = out + out_off[j];
    if (!pl.append && pl.cut0 == pl.cut1) {
        uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
        if (head > len) head = len;
        const uint32_t nwords = (len - head) >> 2;
        const uint8_t *sb = src + head;
        const uint32_t *sa = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(sb) & ~(uintptr_t)3);
        const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(sb) & 3u) << 3;
        uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
        // four rounds of loads are issued before the first store: ~1 KB in flight per warp (the copy is latency bound otherwise)
        for (uint32_t k0 = 0; k0 < nwords; k0 += 128) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + u * 32 + lane;
                lo[u] = k < nwords ? __ldg(sa + k) : 0u;
                hi[u] = (k < nwords && sh) ? __ldg(sa + k + 1) : 0u;   // the last needed byte of word k+1 lies inside the record
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + u * 32 + lane;
                if (k < nwords) dw[k] = __funnelshift_r(lo[u], hi[u], sh);
            }
        }
        if (lane < head) dst[lane] = src[lane];
        const uint32_t done = head + 4u * nwords, tail = len - done;
        if (lane < tail) dst[done + lane] = src[done + lane];
        return;
    }
    const uint32_t r = stream ? stream[j] : (uint32_t)j;
    const uint32_t nl = out_len[j];
    const uint32_t gap = pl.cut1 - pl.cut0;
    const uint32_t body = len - gap;                 // bytes copied from the source
    for (uint32_t k = lane; k < body; k += 32) {
        uint32_t s = k < pl.cut0 || gap == 0 ? k : k + gap;
        uint8_t b = src[s];
        if (pl.append && k < 4) b = (uint8_t)((nl - 4) >> (8 * k));     // block_size of the rewritten record
        dst[k] = b;
    }
    if (pl.append && lane == 0) {
        int32_t sc = score[r];
        uint8_t *t = dst + body;
        t[0] = 'A'; t[1] = 'S'; t[2] = 'i';
        t[3] = (uint8_t)sc; t[4] = (uint8_t)(sc >> 8); t[5] = (uint8_t)(sc >> 16); t[6] = (uint8_t)(sc >> 24);
    }
}

struct InLenU64 { const uint32_t *v; __device__ __forceinline__ unsigned long long operator()(uint64_t i) const { return v[i]; } };

} // namespace msg
