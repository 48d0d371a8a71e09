// decode.cuh -- K1/K2: raw BAM records -> SoA columns + fused filter statistics.
//
// Replaces, per record:  sam_read1's field unpacking (msam_helper.c:267),
// bam_cigar2details (mBamVector.c:23-38), bam_get_summary incl. the MD tokenizer
// (mBamVector.c:40-133), bam_aux_get/bam_aux2i on NM/MD/AS (msam_filter.c:146-162),
// the _FILTER_{L,P,Z} predicates (msam_filter.c:31-35) and the pool-entry decision
// of mFilterFile (msam_filter.c:132-138,181-183).
//
// HBM plan.  A BAM record is [core 36 B | qname | cigar | SEQ | QUAL | aux]; the
// result depends on everything except SEQ/QUAL (~75 % of a PE150 record).  One CTA
// takes 128 consecutive records and stages, per record, two speculative windows with
// coalesced 16-byte streaming loads (LPR = 8 or 16 lanes per record, one load each):
//   head window: `hc` chunks from the record start rounded down to 16 B
//   tail window: `tc` chunks ending at the NEXT record's offset rounded up to 16 B
//                (the aux fields sit at the very end of a record, so the offset index
//                 locates them without reading the record first)
// hc/tc come from a host-side probe of the chunk's typical qname/cigar/aux sizes.  There is
// one dependent load phase (offsets -> windows); SEQ/QUAL sectors are never requested.
// Parsing runs out of shared memory, thread per record, with an odd word stride per slot
// (bank-conflict-free); the MD string is classified four bytes at a time (SWAR) into
// letter/caret bit masks and the reference's tokenizer rule becomes one add + popc.
// Records that do not fit their windows fall back to a byte-wise global-memory parser
// (same template code, other accessor), which `debug_force_slow` exercises in the tests.
#pragma once
#include "common.cuh"

namespace msg {

constexpr int ACCT_SLOTS = 64;                         // accounting counters are spread over this many 128-byte slots
constexpr int DEC_R = 128;                              // records == threads per CTA (64 measured 4 % slower; 256 exceeds static smem at LPR=16)

struct DecodeParams {
    const uint8_t  *raw;
    const uint64_t *off;
    uint64_t n;
    uint64_t nbytes;                // size of the chunk: offsets beyond it are malformed
    uint64_t nbytes_readable;       // bytes the 16-byte window loads may touch (multiple of 16; may be < nbytes for caller-owned buffers)
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
    int32_t  *diff;                 // per-position difference array (depth mode) ...
    unsigned long long *covbits;    // ... or one bit per position + per-target depth sums (summary mode, DM_COV_BITS)
    unsigned long long *covsum;
    const uint64_t *covbase;        // first cell (depth mode) / first 64-bit word (summary mode) of every target
    const uint32_t *tlen;
    uint8_t  *covered;
    int32_t  n_targets;
    uint32_t head_chunks, tail_chunks;   // window split, head_chunks + tail_chunks <= LPR
    // accounting
    uint32_t *err;                  // [0] flags, [1] first offending record (atomicMin)
    unsigned long long *acct;       // ACCT_SLOTS slots of 128 bytes: [0] algorithmic bytes, [1] slow records (summed by the host)
};

// ---------------------------------------------------------------- accessors
struct SmAcc {                       // bytes staged in a shared-memory slot region
    static constexpr bool kSwar = true;
    const uint32_t *w; uint32_t rel;
    __device__ __forceinline__ uint32_t u32(uint32_t x) const {
        uint32_t b = rel + x, i = b >> 2;
        return __funnelshift_r(w[i], w[i + 1], b << 3);          // SHF uses the shift amount mod 32
    }
    __device__ __forceinline__ uint32_t u8(uint32_t x) const {
        uint32_t b = rel + x;
        return (w[b >> 2] >> ((b & 3u) * 8u)) & 0xffu;
    }
};
struct GlAcc {                       // byte-wise global memory (slow path)
    static constexpr bool kSwar = false;
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

// same fields from a staged slot: 6 aligned LDS + 5 funnel shifts instead of 5 unaligned reads
__device__ __forceinline__ RecCore parse_core(const SmAcc &hd, uint64_t rec_len64)
{
    RecCore c;
    const uint32_t *w = hd.w + (hd.rel >> 2);
    const uint32_t sh = hd.rel << 3;
    const uint32_t w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4], w5 = w[5], w6 = w[6];
    c.bad = rec_len64 < 36 || rec_len64 > 0x7fffffffull;
    c.rec_len = (uint32_t)rec_len64;
    c.tid = (int32_t)__funnelshift_r(w1, w2, sh);
    c.pos = (int32_t)__funnelshift_r(w2, w3, sh);
    c.lq  = __funnelshift_r(w3, w4, sh) & 0xffu;
    const uint32_t f4 = __funnelshift_r(w4, w5, sh);
    c.nc = f4 & 0xffffu; c.flag = f4 >> 16;
    c.lseq = (int32_t)__funnelshift_r(w5, w6, sh);
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
        const uint32_t c = hd.u32(x0 + 4 * k);
        const uint32_t bit = 1u << (c & 0xfu); const int32_t w = (int32_t)(c >> 4);
        s.wM     += (bit & 0x0181u) ? w : 0;            // M = X
        s.wI     += (bit & 0x0002u) ? w : 0;            // I
        s.wD     += (bit & 0x0004u) ? w : 0;            // D
        s.wClip  += (bit & 0x0030u) ? w : 0;            // S H
        s.wOther += (bit & 0xfe00u) ? w : 0;            // B and undefined ops: NM path only (mBamVector.c:32-33)
    }
    return s;
}

struct AuxOut { bool hasMD, hasNM, hasAS; int32_t md_letters, nm, as; };

// size of a fixed-width aux value by type byte, 0 for Z/H/B/invalid.  Nibble table over
// 'A'..'P' and 'a'..'p' plus 's'/'S' handled apart keeps this at a handful of instructions.
__device__ __forceinline__ uint32_t aux_fixed_size(uint32_t ty)
{
    // index = ty - 'A' for 'A'..'P' (upper) / ty - 'a' (lower); nibble = size
    //            P O N M L K J I H G F E D C B A
    const unsigned long long UP = 0x0000000400000101ull;   // A=1 C=1 I=4
    //            p o n m l k j i h g f e d c b a
    const unsigned long long LO = 0x0000000400408100ull;   // c=1 d=8 f=4 i=4
    uint32_t r = 0;
    const uint32_t u = ty - 'A', l = ty - 'a';
    if (u < 16) r = (uint32_t)(UP >> (4 * u)) & 0xfu;
    if (l < 16) r = (uint32_t)(LO >> (4 * l)) & 0xfu;
    if ((ty | 0x20u) == 's') r = 2;
    return r;
}

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

// NUL-terminated field starting at y: *zend = index of the NUL.  false if no NUL before `limit`.
template <class A>
__device__ __forceinline__ bool z_skip(const A &ax, uint32_t y, uint32_t limit, uint32_t *zend)
{
    if constexpr (A::kSwar) {
        for (uint32_t z = y; z < limit; z += 4) {
            uint32_t w = ax.u32(z);
            uint32_t nul = ~(((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w | 0x7f7f7f7fu);      // exact zero-byte flags (bit 7 of each byte)
            if (nul) { uint32_t k = (uint32_t)(__ffs((int)nul) - 8) >> 3; if (z + k >= limit) return false; *zend = z + k; return true; }
        }
        return false;
    } else {
        for (uint32_t z = y; z < limit; z++) if (ax.u8(z) == 0) { *zend = z; return true; }
        return false;
    }
}

// MD tokenizer (mBamVector.c:112-118): a maximal run of bytes outside "^0123456789" adds its
// length to `edit` iff it does not start the string and the byte before it is not '^'.
template <class A>
__device__ __forceinline__ bool md_scan_bytes(const A &ax, uint32_t y, uint32_t limit, int32_t *letters_out, uint32_t *zend)
{
    int32_t letters = 0; bool in_run = false, counted = false; uint32_t prev = 0;
    for (uint32_t z = y; z < limit; z++) {
        uint32_t c = ax.u8(z);
        if (c == 0) { *letters_out = letters; *zend = z; return true; }
        bool delim = (c == '^') || (c >= '0' && c <= '9');
        if (delim) in_run = false;
        else {
            if (!in_run) { in_run = true; counted = (z > y) && (prev != '^'); }
            letters += counted ? 1 : 0;
        }
        prev = c;
    }
    return false;
}

__device__ __forceinline__ uint32_t byteflags_to_bits(uint32_t m)
{   // flags at bits 7,15,23,31 -> bits 0..3
    return (((m >> 7) * 0x00204081u) >> 21) & 0xfu;
}

template <class A>
__device__ __forceinline__ bool md_scan(const A &ax, uint32_t y, uint32_t limit, int32_t *letters_out, uint32_t *zend)
{
    if constexpr (!A::kSwar) return md_scan_bytes(ax, y, limit, letters_out, zend);
    else {
        // one bit per character (strings up to 32 chars; longer or non-ASCII -> byte loop)
        uint32_t L = 0, C = 0;
        for (uint32_t pos = 0; pos < 32; pos += 4) {
            if (y + pos >= limit) return false;
            const uint32_t w = ax.u32(y + pos);
            if (w & 0x80808080u) return md_scan_bytes(ax, y, limit, letters_out, zend);
            uint32_t nul = ~(((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w | 0x7f7f7f7fu);
            const uint32_t x = w ^ 0x5e5e5e5eu;                                           // '^'
            uint32_t car = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
            const uint32_t dig = (w + 0x50505050u) & ~(w + 0x46464646u) & 0x80808080u;    // '0'..'9' (bytes < 0x80)
            uint32_t let = ~(nul | car | dig) & 0x80808080u;
            uint32_t k = 4;
            if (nul) {
                k = (uint32_t)(__ffs((int)nul) - 8) >> 3;                                 // first NUL byte in this word
                const uint32_t keep = (1u << (8 * k)) - 1u;
                let &= keep; car &= keep;
            }
            L |= byteflags_to_bits(let) << pos;
            C |= byteflags_to_bits(car) << pos;
            if (nul) {
                if (y + pos + k >= limit) return false;
                // runs that start the string or follow a caret are not counted: adding their lowest bit
                // to L ripples through the run and clears it
                const uint32_t S = ((C << 1) | 1u) & L;
                *letters_out = __popc(L & (L + S));
                *zend = y + pos + k;
                return true;
            }
        }
        return md_scan_bytes(ax, y, limit, letters_out, zend);
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
        const uint32_t vsz = aux_fixed_size(ty);
        if (vsz) {
            if (y + vsz > aux_len) break;
            if (isNM | isAS) {
                int32_t v = aux_int(ty, vsz == 1 ? (w >> 24) : ax.u32(y));   // 1-byte values ride in the header word
                if (isNM) { o.hasNM = true; o.nm = v; }
                if (isAS) { o.hasAS = true; o.as = v; }
            }
            if (isMD) { o.hasMD = true; o.md_letters = 0; }
            y += vsz;
        } else if (ty == 'Z' || ty == 'H') {
            int32_t letters = 0; uint32_t z = y; bool term;
            if (isMD) term = md_scan(ax, y, aux_len, &letters, &z);
            else      term = z_skip(ax, y, aux_len, &z);
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

__device__ __forceinline__ uint32_t hash_step(uint32_t h, uint32_t w) { h = (h ^ w) * 0x9E3779B1u; return h ^ (h >> 15); }
__device__ __forceinline__ uint32_t hash_fin(uint32_t h) { h *= 0x85ebca6bu; return h ^ (h >> 13); }

// QNAME hash (any equal byte strings hash equal) and byte-exact equality with another record's QNAME
// in one pass over the name words.  `other` is only read when do_eq (lengths already equal).
template <class A, class B>
__device__ __forceinline__ uint32_t name_hash_eq(const A &a, uint32_t lq, const B &other, bool do_eq, bool *eq)
{
    uint32_t h = 0x811c9dc5u ^ lq, diff = 0;
    for (uint32_t k = 0; k < lq; k += 4) {
        uint32_t w = a.u32(36 + k);
        const uint32_t rem = lq - k;
        const uint32_t mask = rem < 4 ? (1u << (rem * 8)) - 1u : 0xffffffffu;
        w &= mask;
        h = hash_step(h, w);
        if (do_eq) diff |= (w ^ other.u32(36 + k)) & mask;
    }
    *eq = do_eq && diff == 0;
    return hash_fin(h);
}

// coverage of one alignment: diff-array +1/-1 per M/=/X run (msam_coverage.c:60-86)
template <class A>
__device__ __forceinline__ void cover_record(const A &hd, uint32_t x0, uint32_t nc, int32_t tid, int32_t pos,
                                             int32_t *diff, const uint64_t *covbase, const uint32_t *tlen, uint8_t *covered)
{
    if (!covered[tid]) covered[tid] = 1;                       // :45-49, any record with tid >= 0 (test first, see cover_record_bits)
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

// Summary-only coverage (msam_coverage.c:189-219 needs, per target, "#positions with depth != 0" and "sum of depths"):
// one BIT per reference position instead of a 4-byte difference cell -- a position is touched iff some M/=/X run covers
// it, and the sum of depths is the sum of the (clamped) run lengths.  The bitmap of a 400 Mbp reference set is 50 MB and
// stays in L2, where the diff array (1.6 GB) sends two random read-modify-writes per run to HBM.  Interior words of a
// run are plain stores of all-ones (any interleaving with another run's atomicOr leaves all-ones), the two edge words
// are atomicOr.
__device__ __forceinline__ void cover_bits(unsigned long long *w, long long lo, long long hi)
{
    const long long wlo = lo >> 6, whi = (hi - 1) >> 6;
    const unsigned long long mlo = ~0ull << (lo & 63), mhi = ~0ull >> (63 - ((hi - 1) & 63));
    if (wlo == whi) { atomicOr(w + wlo, mlo & mhi); return; }
    atomicOr(w + wlo, mlo);
    for (long long k = wlo + 1; k < whi; k++) w[k] = ~0ull;
    atomicOr(w + whi, mhi);
}
template <class A>
__device__ __forceinline__ void cover_record_bits(const A &hd, uint32_t x0, uint32_t nc, int32_t tid, int32_t pos,
                                                  unsigned long long *bits, const uint64_t *wbase, const uint32_t *tlen, uint8_t *covered,
                                                  unsigned long long *sum, uint32_t *deferred = nullptr)
{
    if (!covered[tid]) covered[tid] = 1;                       // :45-49, any record with tid >= 0 (test first: tens of millions of
                                                               // stores into the same few cache lines serialise in L2)
    const long long tl = tlen[tid];
    unsigned long long *w = bits + wbase[tid];
    long long q = pos, total = 0;
    for (uint32_t k = 0; k < nc; k++) {
        uint32_t c = hd.u32(x0 + 4 * k);
        uint32_t op = c & 0xfu; long long wd = (long long)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            long long lo = q < 0 ? 0 : q, hi = q + wd > tl ? tl : q + wd;
            if (hi > lo) { cover_bits(w, lo, hi); total += hi - lo; }
            q += wd;
        } else if (op == 2 || op == 3) q += wd;
    }
    // sum of depths += covered bases of this record; with `deferred` the caller adds it (warp-aggregated per target)
    if (deferred && total < (1ll << 26)) *deferred = (uint32_t)total;
    else if (total) atomicAdd(sum + tid, (unsigned long long)total);
}

// one atomic per (warp, target) instead of one per record: lanes with the same target find each other with MATCH.ANY
// and add up with REDUX over the match group.  Called by all 32 lanes at a converged point; tid < 0 or bases == 0:
// nothing to add.  (bases < 2^26 per lane, so the group sum stays inside 32 bits.)
__device__ __forceinline__ void cover_sum_flush(unsigned long long *sum, int32_t tid, uint32_t bases)
{
    const uint32_t lane = threadIdx.x & 31u;
    const bool have = tid >= 0 && bases != 0;
    if (!__any_sync(0xffffffffu, have)) return;
    const uint32_t grp = __match_any_sync(0xffffffffu, have ? tid : -1 - (int32_t)lane);
    const uint32_t tot = __reduce_add_sync(grp, have ? bases : 0u);
    if (have && lane == (uint32_t)__ffs((int)grp) - 1u) atomicAdd(sum + tid, (unsigned long long)tot);
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

// Parse one staged record (thread t of the tile, record index i) and write its SoA row.
// s_off[0..nrec] are the tile's offsets, off_prev the offset of the record before the tile.
template <int SLOT_WORDS>
__device__ __forceinline__ void parse_and_emit(const DecodeParams &p, uint32_t t, uint64_t i, bool active, const uint32_t *s_slot,
                                               const uint64_t *s_off, uint64_t off_prev, uint32_t &alg, uint32_t &nslow, int32_t &cov_tid, uint32_t &cov_bases)
{
    const uint32_t mode = p.mode;
    const bool force_slow = mode & DM_FORCE_SLOW;
    const uint32_t hc = p.head_chunks, tc = p.tail_chunks;
    if (!active) return;
    const uint32_t *slot = s_slot + t * SLOT_WORDS;
    // an index that is not monotonic or runs past the chunk is reported (MSG_EFORMAT) without following it
    const uint64_t o = s_off[t], o1 = s_off[t + 1];
    const bool off_ok = o1 <= p.nbytes && o <= o1;
    const uint64_t len = off_ok ? o1 - o : 0;
    const uint32_t rel = (uint32_t)(o & 15u);
    uint32_t arel = 0;
    bool slow = force_slow || o + 36 > p.nbytes_readable;            // (fixed header not fully staged: chunk's last, tiny record)
    RecCore c;
    if (!off_ok || len < 36) { c = RecCore{-1, 0, 0, 0, 0, 0, 0, 0, 0, true}; slow = false; }
    else c = slow ? parse_core(GlAcc{p.raw + o}, len) : parse_core(SmAcc{slot, rel}, len);
    // does the record fit its windows?
    const uint32_t need_head = 36 + c.lq + ((mode & DM_NEED_CIGAR) ? 4 * c.nc : 0);
    if (rel + need_head > 16 * hc) slow = true;
    // window chunks at or beyond nbytes_readable were not loaded (caller-owned buffers are only read up to the last whole
    // 16-byte chunk): the chunk's final record(s) then take the byte-wise path
    if (o + need_head > p.nbytes_readable || ((mode & DM_NEED_AUX) && c.aux_len && o + c.rec_len > p.nbytes_readable)) slow = true;
    if ((mode & DM_NEED_AUX) && c.aux_len) {
        const uint64_t wend = (o + c.rec_len + 15ull) & ~15ull;           // end of the tail window
        const uint64_t astart = o + c.aux_off;
        if (wend < 16ull * tc || astart < wend - 16ull * tc) slow = true;
        else arel = (uint32_t)(astart - (wend - 16ull * tc));
    }
    RecResult r;
    uint32_t h = 0; bool eq = false;
    if (!c.bad && (mode & DM_NEED_CIGAR) && c.nc == 2 && c.lseq > 0) {
        // htslib stores CIGARs of > 65535 operations in the CG:B,I tag behind the placeholder "<l_seq>S<ref_len>N" (SAM spec
        // 4.2.2) and sam_read1 swaps the real CIGAR in; this parser does not, so it refuses the record instead of mis-counting it
        const GlAcc g0{p.raw + o};
        const uint32_t c0 = g0.u32(36 + c.lq), c1 = g0.u32(36 + c.lq + 4);
        if (c0 == (((uint32_t)c.lseq << 4) | 4u) && (c1 & 0xfu) == 3u) { atomicOr(p.err, DERR_CGTAG); atomicMin(p.err + 1, (uint32_t)min(i, (uint64_t)0xffffffffu)); }
    }
    if (c.bad) {
        r = RecResult{0, 0, 0, 0, 0, c.flag & FB_FLAG_MASK, false};
        atomicOr(p.err, DERR_FORMAT); atomicMin(p.err + 1, (uint32_t)min(i, (uint64_t)0xffffffffu));
    } else if (slow) {
        GlAcc g{p.raw + o};
        GlAcc gx{p.raw + o + c.aux_off};
        r = finish_record(p, c, g, gx);
        const uint64_t op = t ? s_off[t - 1] : off_prev;
        GlAcc gp{p.raw + op};
        h = name_hash_eq(g, c.lq, gp, i > 0 && op + 36 <= o && gp.u8(12) == c.lq, &eq);
        if ((mode & DM_COV_FUSED) && (r.fbits & FB_INPOOL) && c.tid >= 0) {
            if (c.tid >= p.n_targets) atomicOr(p.err, DERR_FORMAT);
            else if (mode & DM_COV_BITS) { cov_tid = c.tid; cover_record_bits(g, 36 + c.lq, c.nc, c.tid, c.pos, p.covbits, p.covbase, p.tlen, p.covered, p.covsum, &cov_bases); }
            else cover_record(g, 36 + c.lq, c.nc, c.tid, c.pos, p.diff, p.covbase, p.tlen, p.covered);
        }
        r.fbits |= FB_SLOW; nslow = 1;
    } else {
        SmAcc hd{slot, rel};
        SmAcc ax{slot + hc * 4, arel};
        r = finish_record(p, c, hd, ax);
        // previous record's QNAME: from its slot when its head window holds the whole name, else from global memory
        bool prev_staged = false; uint32_t prel = 0, plq = 0;
        if (t > 0) {
            prel = (uint32_t)(s_off[t - 1] & 15u);
            SmAcc pv{slot - SLOT_WORDS, prel};
            plq = pv.u32(12) & 0xffu;
            prev_staged = (s_off[t] - s_off[t - 1] >= 36) && (prel + 36 + plq <= 16 * hc);
        }
        if (prev_staged) {
            SmAcc pv{slot - SLOT_WORDS, prel};
            h = name_hash_eq(hd, c.lq, pv, plq == c.lq, &eq);
        } else {
            const uint64_t op = t ? s_off[t - 1] : off_prev;
            GlAcc gp{p.raw + op};
            h = name_hash_eq(hd, c.lq, gp, i > 0 && op + 36 <= o && gp.u8(12) == c.lq, &eq);
        }
        if ((mode & DM_COV_FUSED) && (r.fbits & FB_INPOOL) && c.tid >= 0) {
            if (c.tid >= p.n_targets) atomicOr(p.err, DERR_FORMAT);
            else if (mode & DM_COV_BITS) { cov_tid = c.tid; cover_record_bits(hd, 36 + c.lq, c.nc, c.tid, c.pos, p.covbits, p.covbase, p.tlen, p.covered, p.covsum, &cov_bases); }
            else cover_record(hd, 36 + c.lq, c.nc, c.tid, c.pos, p.diff, p.covbase, p.tlen, p.covered);
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
    uint32_t a = 8u + 36u + c.lq + ((mode & DM_NEED_CIGAR) ? 4u * c.nc : 0u) + ((mode & DM_NEED_AUX) ? c.aux_len : 0u);
    alg += a > (1u << 20) ? (1u << 20) : a;                          // keeps the per-warp sum inside 32 bits
}

// ---------------------------------------------------------------- fast path
// Sequential reader over a staged window: one LDS + one funnel shift per 32-bit word at an arbitrary byte alignment
// (the generic SmAcc::u32 recomputes the word index and loads two words for every access).
struct SeqRd {
    const uint32_t *p; uint32_t sh, cur;
    __device__ __forceinline__ void init(const uint32_t *w, uint32_t byte) { p = w + (byte >> 2); sh = (byte & 3u) << 3; cur = *p; }
    __device__ __forceinline__ uint32_t next() { const uint32_t nx = *++p; const uint32_t v = __funnelshift_r(cur, nx, sh); cur = nx; return v; }
};

// integer aux types c C s S i I: value size in bytes (0 for every other type) and the value itself (bam_aux2i, then the
// reference's int32 truncation) from the 32 bits at the value's position
__device__ __forceinline__ uint32_t aux_int_size(uint32_t ty)
{
    const uint32_t lo = ty | 0x20u;
    return lo == 'c' ? 1u : lo == 's' ? 2u : lo == 'i' ? 4u : 0u;
}
__device__ __forceinline__ int32_t aux_int_value(uint32_t ty, uint32_t sz, uint32_t v)
{
    const uint32_t sh = 32u - 8u * sz;                                   // sz in {1, 2, 4}
    return (ty & 0x20u) ? ((int32_t)(v << sh) >> sh) : (int32_t)((v << sh) >> sh);
}

// The common record, straight-line: windows fit, aux block opens with [NM:<int>] MD:Z:<up to 32 chars> [AS:<int>] (what
// bwa-style aligners and the reference's own generator write, msam validation generator :1034).  Everything is computed
// exactly as finish_record does; anything unusual returns false and the record goes through the generic parser instead.
template <int SLOT_WORDS>
__device__ __forceinline__ bool parse_fast(const DecodeParams &p, uint32_t t, uint64_t i, const uint32_t *s_slot, const uint64_t *s_off,
                                           uint64_t off_prev, uint32_t hc, uint32_t tc, uint32_t &alg, int32_t &cov_tid, uint32_t &cov_bases)
{
    const uint32_t mode = p.mode;
    const uint64_t o = s_off[t], o1 = s_off[t + 1];
    if (!(o1 <= p.nbytes_readable && o <= o1)) return false;           // also covers o1 > nbytes: nbytes_readable <= nbytes + 64 and the generic path re-checks
    const uint64_t len64 = o1 - o;
    if (len64 < 36 || len64 > 0x0fffffffull || o1 > p.nbytes) return false;
    const uint32_t len = (uint32_t)len64, rel = (uint32_t)o & 15u;
    const uint32_t *slot = s_slot + t * SLOT_WORDS;
    // ---- core (6 aligned LDS + 5 funnel shifts)
    const uint32_t *w = slot + (rel >> 2);
    const uint32_t sh = (rel & 3u) << 3;
    const uint32_t w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4], w5 = w[5], w6 = w[6];
    const int32_t tid = (int32_t)__funnelshift_r(w1, w2, sh);
    const int32_t pos = (int32_t)__funnelshift_r(w2, w3, sh);
    const uint32_t lq = __funnelshift_r(w3, w4, sh) & 0xffu;
    const uint32_t f4 = __funnelshift_r(w4, w5, sh);
    const uint32_t nc = f4 & 0xffffu, flag = f4 >> 16;
    const uint32_t lseq = __funnelshift_r(w5, w6, sh);
    if (lseq > 0x07ffffffu) return false;
    const uint32_t aux_off = 36u + lq + 4u * nc + ((lseq + 1u) >> 1) + lseq;
    if (aux_off > len) return false;
    const uint32_t aux_len = len - aux_off;
    const bool need_cigar = mode & DM_NEED_CIGAR, need_aux = mode & DM_NEED_AUX;
    const uint32_t need_head = 36u + lq + (need_cigar ? 4u * nc : 0u);
    if (rel + need_head > 16u * hc) return false;
    if (need_cigar && nc == 2 && lseq) {                           // "<l_seq>S<n>N" announces a CG-tag CIGAR: the generic parser reports it
        SeqRd cr; cr.init(slot, rel + 36u + lq);
        if (cr.next() == ((lseq << 4) | 4u)) return false;
    }
    uint32_t arel = 0;
    if (need_aux && aux_len) {
        const uint64_t wend = (o1 + 15ull) & ~15ull;
        if (wend < 16ull * tc) return false;
        const uint32_t back = (uint32_t)(wend - o1) + aux_len;       // bytes from the start of the aux block to the end of the tail window
        if (back > 16u * tc) return false;
        arel = 16u * tc - back;
    }
    // ---- QNAME hash + exact compare with the previous record's name (both from their slots)
    bool do_eq = false;
    SeqRd pr; pr.p = slot; pr.sh = 0; pr.cur = 0;
    if (t > 0) {
        const uint64_t op = s_off[t - 1];
        const uint32_t prel = (uint32_t)op & 15u;
        const uint32_t *pslot = slot - SLOT_WORDS;
        const uint32_t plq = (pslot[(prel + 12u) >> 2] >> (((prel + 12u) & 3u) << 3)) & 0xffu;
        if (!(o - op >= 36 && prel + 36u + plq <= 16u * hc)) return false;        // previous name not staged: generic path reads it from global memory
        do_eq = plq == lq;
        if (do_eq) pr.init(pslot, prel + 36u);
    }
    uint32_t h = 0x811c9dc5u ^ lq, diff = 0;
    bool eq_glob = false;
    if (t == 0 && i > 0) {
        // first record of the tile: its predecessor lives in another CTA's tile, read that name from global memory
        const GlAcc gp{p.raw + off_prev};
        h = name_hash_eq(SmAcc{slot, rel}, lq, gp, off_prev + 36 <= o && gp.u8(12) == lq, &eq_glob);
    } else {
        SeqRd nr; nr.init(slot, rel + 36u);
        const uint32_t nfull = lq >> 2, remb = lq & 3u;                  // same words and masks as name_hash_eq
        for (uint32_t k = 0; k < nfull; k++) {
            const uint32_t x = nr.next();
            h = hash_step(h, x);
            if (do_eq) diff |= x ^ pr.next();
        }
        if (remb) {
            const uint32_t mask = (1u << (remb * 8)) - 1u;
            const uint32_t x = nr.next() & mask;
            h = hash_step(h, x);
            if (do_eq) diff |= (x ^ pr.next()) & mask;
        }
        h = hash_fin(h);
    }
    const bool eq = (do_eq && diff == 0) || eq_glob;
    // ---- CIGAR
    const bool need_stats = mode & DM_NEED_STATS;
    CigSum cs = {0, 0, 0, 0, 0};
    if (need_stats && nc) {
        SeqRd cr; cr.init(slot, rel + 36u + lq);
        for (uint32_t k = 0; k < nc; k++) {
            const uint32_t c = cr.next();
            const uint32_t bit = 1u << (c & 0xfu); const int32_t wd = (int32_t)(c >> 4);
            cs.wM     += (bit & 0x0181u) ? wd : 0;
            cs.wI     += (bit & 0x0002u) ? wd : 0;
            cs.wD     += (bit & 0x0004u) ? wd : 0;
            cs.wClip  += (bit & 0x0030u) ? wd : 0;
            cs.wOther += (bit & 0xfe00u) ? wd : 0;
        }
    }
    // ---- aux: [NM int] MD:Z [AS int] at the head of the block
    const bool want_as = (p.score != nullptr) && !(mode & DM_RESCORE);
    bool hasMD = false, hasNM = false, hasAS = false; int32_t md_letters = 0, nm = 0, as = 0;
    if (need_aux) {
        const uint32_t *axw = slot + hc * 4;
        uint32_t y = 0;
        if (need_stats) {
            if (aux_len < 4) return false;
            SeqRd ar; ar.init(axw, arel);
            uint32_t a0 = ar.next();
            if ((a0 & 0xffffu) == ('N' | ('M' << 8))) {
                const uint32_t ty = (a0 >> 16) & 0xffu, sz = aux_int_size(ty);
                uint32_t v = a0 >> 24;
                if (sz == 0 || 3u + sz + 4u > aux_len) return false;
                if (sz > 1) { SeqRd vr; vr.init(axw, arel + 3u); v = vr.next(); }
                nm = aux_int_value(ty, sz, v); hasNM = true;
                y = 3u + sz;
                SeqRd a1; a1.init(axw, arel + y); a0 = a1.next();
            }
            if ((a0 & 0xffffffu) != ('M' | ('D' << 8) | ('Z' << 16))) return false;
            // MD string: one bit per character, as md_scan (strings up to 32 chars)
            const uint32_t ms = y + 3u;
            SeqRd mr; mr.init(axw, arel + ms);
            uint32_t L = 0, C = 0, zend = 0; bool done = false;
#pragma unroll 1
            for (uint32_t q = 0; q < 32; q += 4) {
                if (ms + q >= aux_len) return false;
                const uint32_t x = mr.next();
                if (x & 0x80808080u) return false;
                uint32_t nul = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
                const uint32_t xc = x ^ 0x5e5e5e5eu;
                uint32_t car = ~(((xc & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xc | 0x7f7f7f7fu);
                const uint32_t dig = (x + 0x50505050u) & ~(x + 0x46464646u) & 0x80808080u;
                uint32_t let = ~(nul | car | dig) & 0x80808080u;
                if (nul) {
                    const uint32_t kk = (uint32_t)(__ffs((int)nul) - 8) >> 3;
                    const uint32_t keep = (1u << (8 * kk)) - 1u;
                    let &= keep; car &= keep;
                    zend = ms + q + kk; done = true;
                }
                L |= byteflags_to_bits(let) << q;
                C |= byteflags_to_bits(car) << q;
                if (done) break;
            }
            if (!done || zend >= aux_len) return false;
            const uint32_t S = ((C << 1) | 1u) & L;
            md_letters = __popc(L & (L + S));
            hasMD = true;
            y = zend + 1u;
        }
        if (want_as) {
            if (y + 4u > aux_len) return false;
            SeqRd ar; ar.init(axw, arel + y);
            const uint32_t a2 = ar.next();
            if ((a2 & 0xffffu) != ('A' | ('S' << 8))) return false;
            const uint32_t ty = (a2 >> 16) & 0xffu, sz = aux_int_size(ty);
            uint32_t v = a2 >> 24;
            if (sz == 0 || y + 3u + sz > aux_len) return false;
            if (sz > 1) { SeqRd vr; vr.init(axw, arel + y + 3u); v = vr.next(); }
            as = aux_int_value(ty, sz, v); hasAS = true;
        }
    }
    // ---- statistics, filter, outputs: same arithmetic as finish_record
    int32_t alen = 0, qlen = 0, qclip = 0, edit = 0; int path = 0;
    if (need_stats) {
        qclip = cs.wClip; qlen = cs.wM + cs.wI + cs.wClip;
        if (hasMD) { alen = cs.wM + cs.wI + cs.wD; edit = cs.wI + cs.wD + md_letters; path = 2; }
        else if (hasNM) { alen = cs.wM + cs.wI + cs.wD + cs.wOther; edit = nm; path = 1; }
        else { alen = qlen = qclip = edit = 0; }
    }
    const bool mapped = !(flag & BAM_FUNMAP);
    bool has_as = hasAS; int32_t score = as; bool inpool, notag = false;
    if (!(mode & DM_DO_FILTER)) inpool = true;
    else if (!mapped) inpool = (mode & DM_HAS_FILTER) && (mode & DM_KEEP_UNMAP) && p.ppt >= 0 && (mode & DM_INVERT);
    else {
        if ((mode & DM_REQ_STATS) && path == 0) notag = true;
        if (mode & DM_RESCORE) { score = (alen - edit) - edit; has_as = true; }
        if (!(mode & DM_HAS_FILTER)) inpool = true;
        else {
            bool fail = false;
            if (p.min_length > 0 && alen < p.min_length) fail = true;
            if (p.ppt != 0) {
                if (p.ppt < 0) fail |= (1000 * (edit - alen) < alen * p.ppt);
                else           fail |= (1000 * (alen - edit) < alen * p.ppt);
            }
            if (p.max_clip < 100) fail |= (100 * qclip > p.max_clip * qlen);
            inpool = (fail == ((mode & DM_INVERT) != 0));
        }
    }
    uint32_t fbits = (flag & FB_FLAG_MASK) | (inpool ? FB_INPOOL : 0u) | (has_as ? FB_HAS_AS : 0u) | (eq ? FB_EQPREV : 0u);
    if ((mode & DM_COV_FUSED) && inpool && tid >= 0) {
        if (tid >= p.n_targets) atomicOr(p.err, DERR_FORMAT);
        else if (mode & DM_COV_BITS) { cov_tid = tid; cover_record_bits(SmAcc{slot, rel}, 36 + lq, nc, tid, pos, p.covbits, p.covbase, p.tlen, p.covered, p.covsum, &cov_bases); }
        else cover_record(SmAcc{slot, rel}, 36 + lq, nc, tid, pos, p.diff, p.covbase, p.tlen, p.covered);
    }
    if (notag) { atomicOr(p.err, DERR_NOTAG); atomicMin(p.err + 1, (uint32_t)min(i, (uint64_t)0xffffffffu)); }
    if (p.tid)   p.tid[i] = tid;
    if (p.fb)    p.fb[i] = fbits;
    if (p.score) p.score[i] = score;
    if (p.hash)  p.hash[i] = h;
    if (p.alen)  { p.alen[i] = alen; p.qlen[i] = qlen; p.qclip[i] = qclip; p.edit[i] = edit; }
    const uint32_t a = 8u + 36u + lq + (need_cigar ? 4u * nc : 0u) + (need_aux ? aux_len : 0u);
    alg += a > (1u << 20) ? (1u << 20) : a;                          // keeps the per-warp sum inside 32 bits
    return true;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr, bool g64)
{
    if (g64) asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
    else     asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
}

// One CTA per tile of 128 records; ~10 CTAs are resident per SM, which is what hides the two dependent DRAM
// latencies (offsets -> windows) and keeps the issue slots of the (low-ILP) parser busy.  Staging is LDGSTS
// (cp.async.cg, 16 bytes, no register round trip, `.L2::64B` fill granularity): LPR lanes per record, slots are
// 16-byte aligned with one pad chunk (stride LPR*16 + 16 bytes).
template <int LPR, bool G64, bool ASYNC>
__global__ void __launch_bounds__(DEC_R, ASYNC ? 10 : 9) decode_kernel(const __grid_constant__ DecodeParams p)
{
    constexpr int SLOT_WORDS = LPR * 4 + 4;
    __shared__ __align__(16) uint32_t s_slot[DEC_R * SLOT_WORDS + 4];
    __shared__ uint64_t s_off[DEC_R + 1];
    __shared__ uint32_t s_hb[DEC_R], s_tb[DEC_R];

    const uint32_t t = threadIdx.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * DEC_R;
    const uint32_t nrec = (uint32_t)min((uint64_t)DEC_R, p.n - t0);
    const uint64_t i = t0 + t;
    const bool active = t < nrec;
    const uint32_t hc = p.head_chunks, tc = p.tail_chunks;

    // offsets: one coalesced load; each thread also fetches its successor's offset so that the
    // 16-byte chunk indices of both windows are known without a second barrier round
    uint64_t off_prev = 0;
    if (active) {
        const uint64_t o = p.off[i], o1 = p.off[i + 1];
        s_off[t] = o;
        if (t == nrec - 1) s_off[nrec] = o1;
        s_hb[t] = (uint32_t)(o >> 4);                                       // head window: first chunk
        s_tb[t] = (uint32_t)((o1 + 15ull) >> 4) - (hc + tc);                // tail window: chunk index of lane `sub` is s_tb + sub
        if (t == 0 && t0) off_prev = p.off[t0 - 1];
    } else { s_hb[t] = 0xffffffffu - 64u; s_tb[t] = 0xffffffffu - 64u; }    // (fails the bound check below)
    __syncthreads();

    // ---- stage both windows of every record: LPR lanes per record, one 16-byte LDGSTS each
    if (!(p.mode & DM_FORCE_SLOW)) {
        const uint32_t nchunks = (uint32_t)(p.nbytes_readable >> 4);
        const uint32_t sub = t % LPR, r0 = t / LPR;
        if (sub < hc + tc) {
            const uint32_t *wb = (sub < hc ? s_hb : s_tb) + r0;
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_slot + r0 * SLOT_WORDS + sub * 4);
            const uint4 *src = reinterpret_cast<const uint4 *>(p.raw);
            if (ASYNC) {
#pragma unroll
                for (uint32_t it = 0; it < LPR; it++) {
                    const uint32_t cidx = wb[it * (DEC_R / LPR)] + sub;     // a tail window that starts before the buffer wraps to a huge index
                    if (cidx < nchunks) cp_async16(dst + it * (DEC_R / LPR) * SLOT_WORDS * 4, src + cidx, G64);
                }
            } else {
                // register-staged variant (all LPR loads in flight before the first 16-byte shared-memory store): what chunks that
                // live in pinned HOST memory use -- LDGSTS from system memory over PCIe ran at half the rate of plain loads
                uint4 v[LPR]; bool ok[LPR];
#pragma unroll
                for (uint32_t it = 0; it < LPR; it++) {
                    const uint32_t cidx = wb[it * (DEC_R / LPR)] + sub;
                    ok[it] = cidx < nchunks;
                    if (ok[it]) v[it] = G64 ? ldg_stream128(src + cidx) : ldg_stream128_line(src + cidx);
                }
#pragma unroll
                for (uint32_t it = 0; it < LPR; it++)
                    if (ok[it]) *reinterpret_cast<uint4 *>(s_slot + (r0 + it * (DEC_R / LPR)) * SLOT_WORDS + sub * 4) = v[it];
            }
        }
        if (ASYNC) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    }
    __syncthreads();

    uint32_t alg = 0, nslow = 0, cov_bases = 0; int32_t cov_tid = -1;
    if (active) {
        bool done = false;
        if (!(p.mode & DM_FORCE_SLOW)) done = parse_fast<SLOT_WORDS>(p, t, i, s_slot, s_off, off_prev, hc, tc, alg, cov_tid, cov_bases);
        if (!done) parse_and_emit<SLOT_WORDS>(p, t, i, true, s_slot, s_off, off_prev, alg, nslow, cov_tid, cov_bases);
    }
    if (p.mode & DM_COV_BITS) cover_sum_flush(p.covsum, cov_tid, cov_bases);
    alg = __reduce_add_sync(0xffffffffu, alg);         // REDUX: one instruction per warp
    nslow = __reduce_add_sync(0xffffffffu, nslow);
    if ((t & 31u) == 0) {
        // 300 k warps per launch would otherwise add to ONE address (0.7 G same-address atomics/s, not far from what an L2 slice
        // does): 64 slots, 128 bytes apart, picked by CTA
        unsigned long long *slot = p.acct + (size_t)(blockIdx.x & (ACCT_SLOTS - 1)) * 16;
        if (alg) atomicAdd(slot, (unsigned long long)alg);
        if (nslow) atomicAdd(slot + 1, (unsigned long long)nslow);
    }
}

} // namespace msg
