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
    int32_t  *diff;
    const uint64_t *covbase;
    const uint32_t *tlen;
    uint8_t  *covered;
    int32_t  n_targets;
    uint32_t head_chunks, tail_chunks;   // window split, head_chunks + tail_chunks <= LPR
    // accounting
    uint32_t *err;                  // [0] flags, [1] first offending record (atomicMin)
    unsigned long long *acct;       // [0] algorithmic bytes, [1] slow records
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

// Parse one staged record (thread t of the tile, record index i) and write its SoA row.
// s_off[0..nrec] are the tile's offsets, off_prev the offset of the record before the tile.
template <int SLOT_WORDS>
__device__ __forceinline__ void parse_and_emit(const DecodeParams &p, uint32_t t, uint64_t i, bool active, const uint32_t *s_slot,
                                               const uint64_t *s_off, uint64_t off_prev, uint32_t &alg, uint32_t &nslow)
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
            if (c.tid < p.n_targets) cover_record(g, 36 + c.lq, c.nc, c.tid, c.pos, p.diff, p.covbase, p.tlen, p.covered);
            else atomicOr(p.err, DERR_FORMAT);
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
    uint32_t a = 8u + 36u + c.lq + ((mode & DM_NEED_CIGAR) ? 4u * c.nc : 0u) + ((mode & DM_NEED_AUX) ? c.aux_len : 0u);
    alg += a > (1u << 20) ? (1u << 20) : a;                          // keeps the per-warp sum inside 32 bits
}

// One CTA per tile of 128 records; ~11 CTAs are resident per SM, which is what hides the two
// dependent DRAM latencies (offsets -> windows) and keeps the issue slots of the (low-ILP) parser
// busy.  A persistent, register-staged software pipeline was measured and lost (profiles/r01_notes.md):
// at the 4 CTAs/SM its 123 registers allow, the parser alone cannot fill the schedulers.
template <int LPR, bool G64>
__global__ void __launch_bounds__(DEC_R) decode_kernel(const __grid_constant__ DecodeParams p)
{
    constexpr int SLOT_WORDS = LPR * 4 + 1;            // odd stride: thread-per-slot reads are conflict-free
    __shared__ uint32_t s_slot[DEC_R * SLOT_WORDS + 2];
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
    }
    __syncthreads();

    // ---- stage both windows of every record: LPR lanes per record, one 16-byte load each
    if (!(p.mode & DM_FORCE_SLOW)) {
        const uint32_t nchunks = (uint32_t)(p.nbytes_readable >> 4);
        const uint32_t sub = t % LPR, r0 = t / LPR;
        const bool head = sub < hc, go = sub < hc + tc;
        // all LPR loads are issued before the first store so that every thread keeps LPR x 16 B in flight
        uint4 v[LPR];
#pragma unroll
        for (uint32_t it = 0; it < LPR; it++) {
            const uint32_t r = it * (DEC_R / LPR) + r0;
            v[it] = make_uint4(0, 0, 0, 0);
            if (go && r < nrec) {
                const uint32_t cidx = (head ? s_hb[r] : s_tb[r]) + sub;     // a tail window that starts before the buffer wraps to a huge index
                if (cidx < nchunks) v[it] = G64 ? ldg_stream128(reinterpret_cast<const uint4 *>(p.raw) + cidx)
                                                : ldg_stream128_line(reinterpret_cast<const uint4 *>(p.raw) + cidx);
            }
        }
#pragma unroll
        for (uint32_t it = 0; it < LPR; it++) {
            const uint32_t r = it * (DEC_R / LPR) + r0;
            if (go && r < nrec) { uint32_t *d = s_slot + r * SLOT_WORDS + sub * 4; d[0] = v[it].x; d[1] = v[it].y; d[2] = v[it].z; d[3] = v[it].w; }
        }
    }
    __syncthreads();

    uint32_t alg = 0, nslow = 0;
    parse_and_emit<SLOT_WORDS>(p, t, i, active, s_slot, s_off, off_prev, alg, nslow);
    alg = __reduce_add_sync(0xffffffffu, alg);         // REDUX: one instruction per warp
    nslow = __reduce_add_sync(0xffffffffu, nslow);
    if ((t & 31u) == 0) {
        if (alg) atomicAdd(p.acct, (unsigned long long)alg);
        if (nslow) atomicAdd(p.acct + 1, (unsigned long long)nslow);
    }
}

} // namespace msg
