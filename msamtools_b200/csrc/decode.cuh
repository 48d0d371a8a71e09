// decode.cuh -- K1/K2: raw BAM records -> SoA columns + fused filter statistics.
//
// Replaces, per record:  sam_read1's field unpacking (msam_helper.c:267),
// bam_cigar2details (mBamVector.c:23-38), bam_get_summary incl. the MD tokenizer
// (mBamVector.c:40-133), bam_aux_get/bam_aux2i on NM/MD/AS (msam_filter.c:146-162),
// the _FILTER_{L,P,Z} predicates (msam_filter.c:31-35) and the pool-entry decision
// of mFilterFile (msam_filter.c:132-138,181-183).
//
// HBM plan.  A BAM record is [core 36 B | qname | cigar | SEQ | QUAL | aux]; the
// result depends on everything except SEQ/QUAL (~70 % of a PE150 record).  One CTA
// takes 128 consecutive records.  Phase A stages the first 64 B of every record with
// coalesced 16-byte streaming loads (4 lanes per record); each thread then reads its
// record's lengths from shared memory and publishes the exact extra spans it needs
// (rest of qname/cigar, and the aux tail located by the NEXT record's offset); phase B
// fetches those with 8 lanes per record.  SEQ/QUAL sectors are never requested.
// Parsing runs out of shared memory with an odd word stride per record slot so that
// thread-per-record accesses are bank-conflict-free.  Records whose qname/cigar or aux
// exceed the slot fall back to a byte-wise global-memory parser (same code, other
// accessor), which is also what `debug_force_slow` exercises in the tests.
#pragma once
#include "common.cuh"

namespace msg {

constexpr int DEC_R      = 128;                         // records == threads per CTA
constexpr int HEAD_FIX   = 4;                           // 16-B chunks always staged
constexpr int HEAD_X     = 3;                           // optional extra head chunks
constexpr int AUX_C      = 5;                           // aux chunks
constexpr int HEAD_WORDS = (HEAD_FIX + HEAD_X) * 4;     // 28
constexpr int AUX_WORDS  = AUX_C * 4;                   // 20
constexpr int SLOT_WORDS = HEAD_WORDS + AUX_WORDS + 1;  // 49: odd -> conflict-free thread-per-slot
static_assert(HEAD_X + AUX_C == 8, "phase B uses 8 lanes per record");
static_assert((SLOT_WORDS & 1) == 1, "slot stride must be odd");

struct DecodeParams {
    const uint8_t  *raw;
    const uint64_t *off;
    uint64_t n;
    uint64_t nbytes_readable;       // bytes that may be touched (>= round_up(nbytes,16))
    // outputs (any may be null)
    int32_t  *tid;
    uint32_t *fb;
    int32_t  *score;
    uint32_t *hash;
    int32_t  *alen, *qlen, *qclip, *edit;
    // filter
    int32_t  min_length, ppt, max_clip;
    uint32_t mode;
    // fused coverage
    int32_t  *diff;
    const uint64_t *covbase;
    const uint32_t *tlen;
    uint8_t  *covered;
    int32_t  n_targets;
    // accounting
    uint32_t *err;                  // [0] flags, [1] first offending record (atomicMin)
    unsigned long long *acct;       // [0] algorithmic bytes, [1] slow records
};

// ---------------------------------------------------------------- accessors
struct SmAcc {                       // bytes staged in a shared-memory slot region
    const uint32_t *w; uint32_t rel;
    __device__ __forceinline__ uint32_t u32(uint32_t x) const {
        uint32_t b = rel + x, i = b >> 2;
        return __funnelshift_r(w[i], w[i + 1], (b & 3u) * 8u);
    }
    __device__ __forceinline__ uint32_t u8(uint32_t x) const {
        uint32_t b = rel + x;
        return (w[b >> 2] >> ((b & 3u) * 8u)) & 0xffu;
    }
};
struct GlAcc {                       // byte-wise global memory (slow path)
    const uint8_t *p;
    __device__ __forceinline__ uint32_t u8(uint32_t x) const { return p[x]; }
    __device__ __forceinline__ uint32_t u32(uint32_t x) const {
        return (uint32_t)p[x] | ((uint32_t)p[x + 1] << 8) | ((uint32_t)p[x + 2] << 16) | ((uint32_t)p[x + 3] << 24);
    }
};

struct RecCore {
    int32_t  tid, pos, lseq;
    uint32_t lq, nc, flag, rec_len, aux_off, aux_len;
    bool     bad;
};

template <class A>
__device__ __forceinline__ RecCore parse_core(const A &hd, uint64_t rec_len64)
{
    RecCore c;
    c.bad = rec_len64 < 36 || rec_len64 > 0x7fffffffull;
    c.rec_len = (uint32_t)rec_len64;
    c.tid = (int32_t)hd.u32(4);
    c.pos = (int32_t)hd.u32(8);
    c.lq  = hd.u32(12) & 0xffu;
    uint32_t w4 = hd.u32(16);
    c.nc = w4 & 0xffffu; c.flag = w4 >> 16;
    c.lseq = (int32_t)hd.u32(20);
    if (c.lseq < 0) { c.bad = true; c.lseq = 0; }
    uint64_t ao = 36ull + c.lq + 4ull * c.nc + (((uint64_t)c.lseq + 1) >> 1) + (uint64_t)c.lseq;
    if (ao > rec_len64) { c.bad = true; ao = rec_len64; }
    if (c.bad) { c.aux_off = 0; c.aux_len = 0; c.lq = 0; c.nc = 0; }
    else { c.aux_off = (uint32_t)ao; c.aux_len = c.rec_len - c.aux_off; }
    return c;
}

struct CigSum { int32_t wM, wI, wD, wClip, wOther; };

template <class A>
__device__ __forceinline__ CigSum cigar_sum(const A &hd, uint32_t x0, uint32_t nc)
{
    CigSum s = {0, 0, 0, 0, 0};
    for (uint32_t k = 0; k < nc; k++) {
        uint32_t c = hd.u32(x0 + 4 * k);
        uint32_t op = c & 0xfu; int32_t w = (int32_t)(c >> 4);
        if (op == 0 || op == 7 || op == 8) s.wM += w;           // M = X
        else if (op == 1) s.wI += w;                            // I
        else if (op == 2) s.wD += w;                            // D
        else if (op == 4 || op == 5) s.wClip += w;              // S H
        else if (op != 3 && op != 6) s.wOther += w;             // B and undefined ops: NM path only (mBamVector.c:32-33)
    }
    return s;
}

struct AuxOut { bool hasMD, hasNM, hasAS; int32_t md_letters, nm, as; };

__device__ __forceinline__ int32_t aux_int(uint32_t ty, uint32_t v)
{   // htslib bam_aux2i then the reference's (int32_t) truncation
    switch (ty) {
    case 'c': return (int32_t)(int8_t)(v & 0xff);
    case 'C': return (int32_t)(v & 0xff);
    case 's': return (int32_t)(int16_t)(v & 0xffff);
    case 'S': return (int32_t)(v & 0xffff);
    case 'i': case 'I': return (int32_t)v;
    default:  return 0;
    }
}

template <class A>
__device__ __forceinline__ AuxOut aux_scan(const A &ax, uint32_t aux_len, bool need_mdnm, bool need_as)
{
    AuxOut o = {false, false, false, 0, 0, 0};
    const uint32_t TAG_MD = 'M' | ('D' << 8), TAG_NM = 'N' | ('M' << 8), TAG_AS = 'A' | ('S' << 8);
    uint32_t y = 0;
    while (y + 3 <= aux_len) {
        uint32_t w = ax.u32(y);
        uint32_t tag = w & 0xffffu, ty = (w >> 16) & 0xffu;
        y += 3;
        bool isMD = need_mdnm && tag == TAG_MD && !o.hasMD;
        bool isNM = need_mdnm && tag == TAG_NM && !o.hasNM;
        bool isAS = need_as && tag == TAG_AS && !o.hasAS;
        uint32_t vsz = 0;
        if (ty == 'A' || ty == 'c' || ty == 'C') vsz = 1;
        else if (ty == 's' || ty == 'S') vsz = 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') vsz = 4;
        else if (ty == 'd') vsz = 8;
        if (vsz) {
            if (y + vsz > aux_len) break;
            if (isNM | isAS) {
                int32_t v = aux_int(ty, ax.u32(y));
                if (isNM) { o.hasNM = true; o.nm = v; }
                if (isAS) { o.hasAS = true; o.as = v; }
            }
            if (isMD) { o.hasMD = true; o.md_letters = 0; }
            y += vsz;
        } else if (ty == 'Z' || ty == 'H') {
            // MD tokenizer (mBamVector.c:112-118): a maximal run of bytes outside
            // "^0123456789" adds its length iff it does not start the string and
            // the byte before it is not '^'.
            int32_t letters = 0; bool in_run = false, counted = false; uint32_t prev = 0;
            uint32_t z = y; bool term = false;
            while (z < aux_len) {
                uint32_t c = ax.u8(z);
                if (c == 0) { term = true; break; }
                bool delim = (c == '^') || (c >= '0' && c <= '9');
                if (delim) in_run = false;
                else {
                    if (!in_run) { in_run = true; counted = (z > y) && (prev != '^'); }
                    letters += counted ? 1 : 0;
                }
                prev = c; z++;
            }
            if (!term) break;
            if (isMD) { o.hasMD = true; o.md_letters = letters; }
            if (isNM) { o.hasNM = true; o.nm = 0; }
            if (isAS) { o.hasAS = true; o.as = 0; }
            y = z + 1;
        } else if (ty == 'B') {
            if (y + 5 > aux_len) break;
            uint32_t st = ax.u8(y), cnt = ax.u32(y + 1), es;
            if (st == 'c' || st == 'C') es = 1; else if (st == 's' || st == 'S') es = 2;
            else if (st == 'i' || st == 'I' || st == 'f') es = 4; else break;
            unsigned long long tot = 5ull + (unsigned long long)cnt * es;
            if ((unsigned long long)y + tot > aux_len) break;
            if (isMD) { o.hasMD = true; o.md_letters = 0; }
            if (isNM) { o.hasNM = true; o.nm = 0; }
            if (isAS) { o.hasAS = true; o.as = 0; }
            y += (uint32_t)tot;
        } else break;
        if ((!need_mdnm || o.hasMD) && (!need_as || o.hasAS)) break;   // nothing left to find
    }
    return o;
}

template <class A>
__device__ __forceinline__ uint32_t name_hash(const A &hd, uint32_t lq)
{
    uint32_t h = 0x811c9dc5u ^ lq;
    for (uint32_t k = 0; k < lq; k += 4) {
        uint32_t w = hd.u32(36 + k);
        uint32_t rem = lq - k;
        if (rem < 4) w &= (1u << (rem * 8)) - 1u;
        h = (h ^ w) * 0x9E3779B1u;
        h ^= h >> 15;
    }
    h *= 0x85ebca6bu; h ^= h >> 13;
    return h;
}

template <class A, class B>
__device__ __forceinline__ bool name_equal(const A &a, uint32_t lqa, const B &b, uint32_t lqb)
{
    if (lqa != lqb) return false;
    for (uint32_t k = 0; k < lqa; k += 4) {
        uint32_t x = a.u32(36 + k) ^ b.u32(36 + k);
        uint32_t rem = lqa - k;
        if (rem < 4) x &= (1u << (rem * 8)) - 1u;
        if (x) return false;
    }
    return true;
}

// coverage of one alignment: diff-array +1/-1 per M/=/X run (msam_coverage.c:60-86)
template <class A>
__device__ __forceinline__ void cover_record(const A &hd, uint32_t x0, uint32_t nc, int32_t tid, int32_t pos,
                                             int32_t *diff, const uint64_t *covbase, const uint32_t *tlen, uint8_t *covered)
{
    covered[tid] = 1;                                          // :45-49, any record with tid >= 0
    const long long tl = tlen[tid];
    int32_t *d = diff + covbase[tid];
    long long q = pos;
    for (uint32_t k = 0; k < nc; k++) {
        uint32_t c = hd.u32(x0 + 4 * k);
        uint32_t op = c & 0xfu; long long w = (long long)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            long long lo = q < 0 ? 0 : q, hi = q + w > tl ? tl : q + w;
            if (hi > lo) { atomicAdd(d + lo, 1); atomicAdd(d + hi, -1); }   // slot tl is the per-target spill cell
            q += w;
        } else if (op == 2 || op == 3) q += w;
    }
}

// everything after the core: qname hash/compare inputs, cigar, aux, filter, outputs
struct RecResult { int32_t alen, qlen, qclip, edit, score; uint32_t fbits; bool notag; };

template <class A, class X>
__device__ __forceinline__ RecResult finish_record(const DecodeParams &p, const RecCore &c, const A &hd, const X &ax)
{
    RecResult r = {0, 0, 0, 0, 0, 0, false};
    const uint32_t mode = p.mode;
    const bool mapped = !(c.flag & BAM_FUNMAP);
    const bool need_stats = mode & DM_NEED_STATS;
    const bool want_as = (p.score != nullptr) && !(mode & DM_RESCORE);   // host allocates score[] iff AS matters
    AuxOut a = {false, false, false, 0, 0, 0};
    if (mode & DM_NEED_AUX) a = aux_scan(ax, c.aux_len, need_stats, want_as);
    int path = 0;
    if (need_stats) {
        CigSum s = cigar_sum(hd, 36 + c.lq, c.nc);
        r.qclip = s.wClip;
        r.qlen = s.wM + s.wI + s.wClip;
        if (a.hasMD) { r.alen = s.wM + s.wI + s.wD; r.edit = s.wI + s.wD + a.md_letters; path = 2; }        // bam_get_summary
        else if (a.hasNM) { r.alen = s.wM + s.wI + s.wD + s.wOther; r.edit = a.nm; path = 1; }               // bam_cigar2details + NM
        else { r.alen = r.qlen = r.qclip = r.edit = 0; }
    }
    bool has_as = a.hasAS; r.score = a.as;
    bool inpool;
    if (!(mode & DM_DO_FILTER)) inpool = true;
    else if (!mapped) {
        inpool = (mode & DM_HAS_FILTER) && (mode & DM_KEEP_UNMAP) && p.ppt >= 0 && (mode & DM_INVERT);       // :132-136
    } else {
        if ((mode & DM_REQ_STATS) && path == 0) r.notag = true;                                                         // :150-152
        if (mode & DM_RESCORE) { r.score = (r.alen - r.edit) - r.edit; has_as = true; }                      // :160-168
        if (!(mode & DM_HAS_FILTER)) inpool = true;
        else {
            // int32 arithmetic exactly as the reference macros (wraps identically on overflow)
            bool fail = false;
            if (p.min_length > 0 && r.alen < p.min_length) fail = true;
            if (p.ppt != 0) {
                if (p.ppt < 0) fail |= (1000 * (r.edit - r.alen) < r.alen * p.ppt);
                else           fail |= (1000 * (r.alen - r.edit) < r.alen * p.ppt);
            }
            if (p.max_clip < 100) fail |= (100 * r.qclip > p.max_clip * r.qlen);
            inpool = (fail == ((mode & DM_INVERT) != 0));                                                    // :181
        }
    }
    r.fbits = (c.flag & FB_FLAG_MASK) | (inpool ? FB_INPOOL : 0u) | (has_as ? FB_HAS_AS : 0u);
    return r;
}

__global__ void __launch_bounds__(DEC_R) decode_kernel(const __grid_constant__ DecodeParams p)
{
    __shared__ uint32_t s_slot[DEC_R * SLOT_WORDS + 2];
    __shared__ uint64_t s_off[DEC_R + 1];
    __shared__ uint64_t s_auxa[DEC_R];
    __shared__ uint64_t s_offprev;
    __shared__ uint8_t  s_xh[DEC_R], s_auxn[DEC_R], s_slow[DEC_R], s_lq[DEC_R];

    const uint32_t t = threadIdx.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * DEC_R;
    const uint32_t nrec = (uint32_t)min((uint64_t)DEC_R, p.n - t0);
    const uint64_t i = t0 + t;
    const bool active = t < nrec;
    const uint32_t mode = p.mode;
    const bool force_slow = mode & DM_FORCE_SLOW;

    if (active) s_off[t] = p.off[i];
    if (t == 0) { s_off[nrec] = p.off[t0 + nrec]; s_offprev = t0 ? p.off[t0 - 1] : 0; }
    __syncthreads();

    // ---- phase A: first HEAD_FIX chunks of every record, 4 lanes per record
    if (!force_slow) {
        for (uint32_t g = t; g < nrec * HEAD_FIX; g += DEC_R) {
            uint32_t r = g / HEAD_FIX, cidx = g % HEAD_FIX;
            uint64_t a = (s_off[r] & ~15ull) + 16ull * cidx;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (a + 16 <= p.nbytes_readable) v = ldg_stream128(reinterpret_cast<const uint4 *>(p.raw + a));
            uint32_t *d = s_slot + r * SLOT_WORDS + cidx * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();

    // ---- core fields, extents of the extra spans
    RecCore c; c.bad = false; c.lq = 0; c.nc = 0; c.aux_len = 0; c.aux_off = 0; c.flag = 0; c.tid = -1; c.pos = 0; c.lseq = 0; c.rec_len = 0;
    bool slow = force_slow;
    uint32_t rel = 0, arel = 0;
    const uint32_t *slot = s_slot + t * SLOT_WORDS;
    if (active) {
        const uint64_t o = s_off[t], len = s_off[t + 1] - o;
        rel = (uint32_t)(o & 15u);
        if (force_slow) c = parse_core(GlAcc{p.raw + o}, len);
        else            c = parse_core(SmAcc{slot, rel}, len);
        uint32_t need_head = 36 + c.lq + ((mode & DM_NEED_CIGAR) ? 4 * c.nc : 0);
        uint32_t hchunks = (rel + need_head + 15) >> 4;
        uint32_t achunks = 0; uint64_t a0 = 0;
        if ((mode & DM_NEED_AUX) && c.aux_len) {
            a0 = (o + c.aux_off) & ~15ull;
            achunks = (uint32_t)((((o + c.rec_len + 15) & ~15ull) - a0) >> 4);
            arel = (uint32_t)((o + c.aux_off) & 15u);
        }
        if (hchunks > HEAD_FIX + HEAD_X || achunks > AUX_C) slow = true;
        s_xh[t]   = slow ? 0 : (uint8_t)(hchunks > HEAD_FIX ? hchunks - HEAD_FIX : 0);
        s_auxn[t] = slow ? 0 : (uint8_t)achunks;
        s_auxa[t] = a0;
        s_slow[t] = slow;
        s_lq[t]   = (uint8_t)c.lq;
    }
    __syncthreads();

    // ---- phase B: exact extra spans, 8 lanes per record (3 head + 5 aux)
    if (!force_slow) {
        const uint32_t lane = t & 31u, wbase = t & ~31u;
#pragma unroll
        for (uint32_t it = 0; it < 8; it++) {
            uint32_t r = wbase + it * 4 + (lane >> 3), sub = lane & 7u;
            if (r < nrec) {
                uint64_t a; uint32_t *d; bool go;
                if (sub < HEAD_X) {
                    go = sub < s_xh[r];
                    a = (s_off[r] & ~15ull) + 16ull * (HEAD_FIX + sub);
                    d = s_slot + r * SLOT_WORDS + (HEAD_FIX + sub) * 4;
                } else {
                    uint32_t cidx = sub - HEAD_X;
                    go = cidx < s_auxn[r];
                    a = s_auxa[r] + 16ull * cidx;
                    d = s_slot + r * SLOT_WORDS + HEAD_WORDS + cidx * 4;
                }
                if (go) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (a + 16 <= p.nbytes_readable) v = ldg_stream128(reinterpret_cast<const uint4 *>(p.raw + a));
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        }
    }
    __syncthreads();

    unsigned long long alg = 0;
    uint32_t nslow = 0;
    if (active) {
        const uint64_t o = s_off[t];
        RecResult r;
        uint32_t h = 0; bool eq = false;
        if (c.bad) {
            r = RecResult{0, 0, 0, 0, 0, c.flag & FB_FLAG_MASK, false};
            atomicOr(p.err, DERR_FORMAT); atomicMin(p.err + 1, (uint32_t)min(i, (uint64_t)0xffffffffu));
        } else if (slow) {
            GlAcc g{p.raw + o};
            GlAcc gx{p.raw + o + c.aux_off};
            r = finish_record(p, c, g, gx);
            if (mode & DM_WANT_HASH) h = name_hash(g, c.lq);
            if (i > 0) {
                uint64_t op = t ? s_off[t - 1] : s_offprev;
                GlAcc gp{p.raw + op};
                eq = name_equal(g, c.lq, gp, gp.u8(12));
            }
            if ((mode & DM_COV_FUSED) && (r.fbits & FB_INPOOL) && c.tid >= 0) {
                if (c.tid < p.n_targets) cover_record(g, 36 + c.lq, c.nc, c.tid, c.pos, p.diff, p.covbase, p.tlen, p.covered);
                else atomicOr(p.err, DERR_FORMAT);
            }
            r.fbits |= FB_SLOW; nslow = 1;
        } else {
            SmAcc hd{slot, rel};
            SmAcc ax{slot + HEAD_WORDS, arel};
            r = finish_record(p, c, hd, ax);
            if (mode & DM_WANT_HASH) h = name_hash(hd, c.lq);
            if (i > 0) {
                if (t > 0 && !s_slow[t - 1]) {
                    SmAcc pv{slot - SLOT_WORDS, (uint32_t)(s_off[t - 1] & 15u)};
                    eq = name_equal(hd, c.lq, pv, s_lq[t - 1]);
                } else {
                    uint64_t op = t ? s_off[t - 1] : s_offprev;
                    GlAcc gp{p.raw + op};
                    eq = name_equal(hd, c.lq, gp, gp.u8(12));
                }
            }
            if ((mode & DM_COV_FUSED) && (r.fbits & FB_INPOOL) && c.tid >= 0) {
                if (c.tid < p.n_targets) cover_record(hd, 36 + c.lq, c.nc, c.tid, c.pos, p.diff, p.covbase, p.tlen, p.covered);
                else atomicOr(p.err, DERR_FORMAT);
            }
        }
        if (r.notag) { atomicOr(p.err, DERR_NOTAG); atomicMin(p.err + 1, (uint32_t)min(i, (uint64_t)0xffffffffu)); }
        if (eq) r.fbits |= FB_EQPREV;
        if (p.tid)   p.tid[i] = c.tid;
        if (p.fb)    p.fb[i] = r.fbits;
        if (p.score) p.score[i] = r.score;
        if (p.hash)  p.hash[i] = h;
        if (p.alen)  { p.alen[i] = r.alen; p.qlen[i] = r.qlen; p.qclip[i] = r.qclip; p.edit[i] = r.edit; }
        // algorithmic bytes A(rec) = 8 (index) + 36 + qname [+ cigar] [+ aux]   
        alg = 8ull + 36ull + c.lq + ((mode & DM_NEED_CIGAR) ? 4ull * c.nc : 0ull) + ((mode & DM_NEED_AUX) ? c.aux_len : 0u);
    }
    alg = warp_sum_u64(alg);
    nslow = warp_sum_u32(nslow);
    if ((t & 31u) == 0) {
        if (alg) atomicAdd(p.acct, alg);
        if (nslow) atomicAdd(p.acct + 1, (unsigned long long)nslow);
    }
}

} // namespace msg
