/*
 * finflate.c -- a fast one-shot raw-DEFLATE (RFC 1951) decoder for BGZF block payloads.
 *
 * BGZF blocks are small (<= 64 KiB), independent and carry their inflated size and CRC32, so a
 * decoder can be one-shot and strict: decode `in_len` bytes into exactly `out_len` bytes or say
 * "no" -- the caller (bamio.c) then hands the block to zlib, and checks the CRC either way.  What
 * makes it faster than a streaming inflate: a 64-bit bit buffer refilled eight bytes at a time
 * (one branch-free refill per symbol: >= 56 bits cover a length code with its extra bits AND the
 * distance code with its extra bits), two-level lookup tables (11 / 8 primary bits) whose entries
 * hold the base value and the TOTAL number of bits to consume (code + extra bits: one shift per
 * symbol, the extra bits are cut out of the saved buffer), the table entry of the NEXT symbol
 * loaded before the current match is copied (so the literal-or-match branch, which is close to a
 * coin flip on BAM data, resolves as soon as it is reached), and 8-byte match copies.  Never writes
 * outside [out, out + out_len): neighbouring blocks are inflated by other threads.
 */
#include "finflate.h"
#include <string.h>

#define LIT_TBITS 11
#define DST_TBITS 8
#define LIT_TABSZ 8192   /* 2^11 + at most 286 subtables of <= 16 entries */
#define DST_TABSZ 4352   /* 2^8  + at most 30 subtables of <= 128 entries */

/* table entry: [0,8) bits to consume at this table level = code bits + extra bits (0: invalid) | [8,12) the code bits among them
 * (link entries: subtable index bits) | 13 literal | 14 end of block | 15 link | [16,32) value (literal, base, subtable base) */
#define E_LITERAL (1u << 13)
#define E_EOB     (1u << 14)
#define E_LINK    (1u << 15)

static const uint16_t LEN_BASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t  LEN_XBITS[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t DST_BASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t  DST_XBITS[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};

static inline uint64_t le64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }   /* little-endian hosts only (x86-64, aarch64) */

static uint32_t bitrev(uint32_t code, int len)
{
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

/* what a decoded symbol means, as table-entry payload (without the bit count) */
static uint32_t sym_entry(int kind, int sym)
{
    if (kind == 0) {                                   /* literal / length alphabet */
        if (sym < 256) return E_LITERAL | ((uint32_t)sym << 16);
        if (sym == 256) return E_EOB;
        if (sym <= 285) return (uint32_t)LEN_BASE[sym - 257] << 16;
        return 0xffffffffu;                            /* 286, 287: never valid in a stream */
    }
    if (kind == 1) {                                   /* distance alphabet */
        if (sym < 30) return (uint32_t)DST_BASE[sym] << 16;
        return 0xffffffffu;
    }
    return (uint32_t)sym << 16;                        /* code-length alphabet: plain symbol */
}

static uint32_t sym_xbits(int kind, int sym)
{
    if (kind == 0) return sym >= 257 && sym <= 285 ? LEN_XBITS[sym - 257] : 0u;
    if (kind == 1) return sym < 30 ? DST_XBITS[sym] : 0u;
    return 0u;
}

/* Canonical Huffman code -> two-level table.  Returns 0, or -1 for anything but a complete code
 * (over-subscribed; incomplete or empty unless it is the distance alphabet): the caller falls back to zlib for those. */
static int build_table(const uint8_t *lens, int nsym, int kind, int tbits, uint32_t *tab, int tabsz)
{
    int count[16] = {0}, maxlen = 0;
    for (int s = 0; s < nsym; s++) count[lens[s]]++;
    count[0] = 0;
    for (int l = 1; l <= 15; l++) if (count[l]) maxlen = l;
    const uint32_t psize0 = 1u << tbits;
    if (!maxlen) {                                     /* no codes at all: fine for the distance alphabet of a literal-only block */
        if (kind != 1) return -1;
        memset(tab, 0, psize0 * sizeof *tab);
        return 0;
    }
    long left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return -1; }
    /* incomplete codes: only the one zlib accepts too -- a distance alphabet with a single one-bit code (deflate writers
     * emit it for blocks with one distinct distance); its unassigned slot stays invalid */
    if (left != 0 && (kind != 1 || maxlen != 1)) return -1;
    uint32_t next[16]; uint32_t code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const uint32_t psize = 1u << tbits;
    memset(tab, 0, psize * sizeof *tab);               /* 0 bits to drop = invalid entry */
    /* pass 1: short codes into the primary table; longest code behind every primary slot that needs a subtable */
    uint8_t sublen[1 << LIT_TBITS];
    if (maxlen > tbits) memset(sublen, 0, psize);
    uint32_t codes[288];
    for (int s = 0; s < nsym; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = bitrev(next[l]++, l);
        codes[s] = r;
        if (l <= tbits) {
            const uint32_t e = sym_entry(kind, s);
            if (e == 0xffffffffu) continue;            /* unusable symbol: its slots stay invalid */
            const uint32_t ent = e | ((uint32_t)l + sym_xbits(kind, s)) | ((uint32_t)l << 8);
            for (uint32_t k = r; k < psize; k += 1u << l) tab[k] = ent;
        } else {
            const uint32_t slot = r & (psize - 1);
            if (l - tbits > sublen[slot]) sublen[slot] = (uint8_t)(l - tbits);
        }
    }
    if (maxlen <= tbits) return 0;
    /* pass 2: allocate the subtables, then fill them */
    uint32_t used = psize;
    for (uint32_t slot = 0; slot < psize; slot++) {
        if (!sublen[slot]) continue;
        const uint32_t sz = 1u << sublen[slot];
        if (used + sz > (uint32_t)tabsz || used > 0xffffu) return -1;
        tab[slot] = E_LINK | ((uint32_t)sublen[slot] << 8) | (used << 16) | (uint32_t)tbits;
        memset(tab + used, 0, sz * sizeof *tab);
        used += sz;
    }
    for (int s = 0; s < nsym; s++) {
        const int l = lens[s];
        if (l <= tbits) continue;
        const uint32_t e = sym_entry(kind, s);
        if (e == 0xffffffffu) continue;
        const uint32_t r = codes[s], link = tab[r & (psize - 1)];
        const uint32_t base = link >> 16, sbits = (link >> 8) & 15u, rem = (uint32_t)(l - tbits);
        const uint32_t ent = e | (rem + sym_xbits(kind, s)) | (rem << 8);
        for (uint32_t k = r >> tbits; k < (1u << sbits); k += 1u << rem) tab[base + k] = ent;
    }
    return 0;
}

int fi_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    const uint8_t *ip = in, *const in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_len;
    uint64_t bitbuf = 0; unsigned bitcnt = 0;
    uint32_t lit_tab[LIT_TABSZ], dst_tab[DST_TABSZ];

    /* top up to >= 56 valid bits (fewer only at the very end of the input; missing bits read as zeros) */
#define REFILL() do { \
        if (in_end - ip >= 8) { bitbuf |= le64(ip) << bitcnt; ip += (63u - bitcnt) >> 3; bitcnt |= 56u; } \
        else while (bitcnt <= 56u && ip < in_end) { bitbuf |= (uint64_t)*ip++ << bitcnt; bitcnt += 8u; } \
    } while (0)
#define DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)
#define BITS(n) ((uint32_t)bitbuf & ((1u << (n)) - 1u))
    /* every DROP is preceded by a check that the bits were really there: `avail` tracks underflow at the end of input */
#define NEED(n) do { if (bitcnt < (unsigned)(n)) return -1; } while (0)

    for (;;) {
        REFILL();
        NEED(3);
        const uint32_t bfinal = BITS(1), btype = (uint32_t)(bitbuf >> 1) & 3u;
        DROP(3);
        if (btype == 0) {                              /* stored */
            DROP(bitcnt & 7u);                         /* to the byte boundary */
            ip -= bitcnt >> 3; bitbuf = 0; bitcnt = 0; /* give the whole bytes still in the buffer back */
            if (in_end - ip < 4) return -1;
            const uint32_t len = ip[0] | (uint32_t)ip[1] << 8, nlen = ip[2] | (uint32_t)ip[3] << 8;
            ip += 4;
            if ((len ^ nlen) != 0xffffu || (size_t)(in_end - ip) < len || (size_t)(out_end - op) < len) return -1;
            memcpy(op, ip, len); op += len; ip += len;
        } else if (btype == 3) {
            return -1;
        } else {
            if (btype == 1) {                          /* fixed code */
                uint8_t lens[288];
                memset(lens, 8, 144); memset(lens + 144, 9, 112); memset(lens + 256, 7, 24); memset(lens + 280, 8, 8);
                if (build_table(lens, 288, 0, LIT_TBITS, lit_tab, LIT_TABSZ)) return -1;
                memset(lens, 5, 32);
                if (build_table(lens, 32, 1, DST_TBITS, dst_tab, DST_TABSZ)) return -1;
            } else {                                   /* dynamic code */
                static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                REFILL(); NEED(14);
                const uint32_t hlit = BITS(5) + 257; DROP(5);
                const uint32_t hdist = BITS(5) + 1; DROP(5);
                const uint32_t hclen = BITS(4) + 4; DROP(4);
                if (hlit > 286 || hdist > 30) return -1;
                uint8_t cl[19] = {0};
                for (uint32_t k = 0; k < hclen; k++) { REFILL(); NEED(3); cl[ORDER[k]] = (uint8_t)BITS(3); DROP(3); }
                uint32_t pre_tab[128];
                if (build_table(cl, 19, 2, 7, pre_tab, 128)) return -1;
                uint8_t lens[286 + 30 + 138];
                uint32_t n = 0;
                while (n < hlit + hdist) {
                    REFILL();
                    const uint32_t e = pre_tab[BITS(7)];
                    if (!(e & 0xffu)) return -1;
                    NEED(e & 0xffu); DROP(e & 0xffu);
                    const uint32_t sym = e >> 16;
                    if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) { if (!n) return -1; NEED(2); rep = 3 + BITS(2); DROP(2); val = lens[n - 1]; }
                    else if (sym == 17) { NEED(3); rep = 3 + BITS(3); DROP(3); }
                    else { NEED(7); rep = 11 + BITS(7); DROP(7); }
                    if (n + rep > hlit + hdist) return -1;
                    memset(lens + n, (int)val, rep); n += rep;
                }
                if (!lens[256]) return -1;             /* no end-of-block code */
                if (build_table(lens, (int)hlit, 0, LIT_TBITS, lit_tab, LIT_TABSZ)) return -1;
                if (build_table(lens + hlit, (int)hdist, 1, DST_TBITS, dst_tab, DST_TABSZ)) return -1;
            }
            /* ---- symbols */
            /* value of a length / distance symbol: base + the extra bits, which sit in the saved buffer right behind the code bits */
#define XVAL(e, saved) (((e) >> 16) + ((uint32_t)((saved) >> (((e) >> 8) & 15u)) & ((1u << (((e) & 0xffu) - (((e) >> 8) & 15u))) - 1u)))
            for (;;) {
                /* Fast iterations while both buffers have slack.  Invariant at the top: the buffer was just refilled (>= 56 bits) and
                 * `e` is the primary-table entry of the symbol that starts at its low end.  One symbol per iteration -- a literal
                 * (<= 15 bits) or a whole match (length <= 15 + 5, distance <= 15 + 13: 48 bits) -- then ONE refill (8 bytes read,
                 * <= 7 consumed) and the lookup for the next symbol, issued before the match is copied.  in_end - ip >= 16 covers
                 * the refill, out_end - op >= 320 a literal or a 258-byte match with its <= 7-byte copy overshoot: no
                 * availability or bounds checks in here. */
                if (in_end - ip >= 16 && out_end - op >= 320) {
                    bitbuf |= le64(ip) << bitcnt; ip += (63u - bitcnt) >> 3; bitcnt |= 56u;
                    uint32_t e = lit_tab[BITS(LIT_TBITS)];
                    do {
                        if (e & E_LINK) { DROP(LIT_TBITS); e = lit_tab[(e >> 16) + BITS((e >> 8) & 15u)]; }
                        const uint64_t saved = bitbuf;
                        DROP(e & 0xffu);
                        if (e & E_LITERAL) {
                            const uint8_t lit = (uint8_t)(e >> 16);
                            e = lit_tab[BITS(LIT_TBITS)];                  /* >= 41 bits are left: look up first, refill behind it (a refill */
                            bitbuf |= le64(ip) << bitcnt; ip += (63u - bitcnt) >> 3; bitcnt |= 56u;   /* only adds bits above the valid ones) */
                            *op++ = lit;
                            continue;
                        }
                        if (!(e & 0xffu)) return -1;
                        if (e & E_EOB) goto block_done;
                        const uint32_t len = XVAL(e, saved);
                        uint32_t d = dst_tab[BITS(DST_TBITS)];
                        if (d & E_LINK) { DROP(DST_TBITS); d = dst_tab[(d >> 16) + BITS((d >> 8) & 15u)]; }
                        if (!(d & 0xffu)) return -1;
                        const uint32_t dist = XVAL(d, bitbuf);
                        DROP(d & 0xffu);
                        bitbuf |= le64(ip) << bitcnt; ip += (63u - bitcnt) >> 3; bitcnt |= 56u;
                        e = lit_tab[BITS(LIT_TBITS)];                      /* next symbol's entry is on its way while the match is copied */
                        if (dist > (size_t)(op - out)) return -1;
                        const uint8_t *src = op - dist;
                        uint8_t *const stop = op + len;
                        if (dist >= 8) {
                            /* two words unconditionally (most matches are shorter than 16 bytes: no data-dependent loop exit to
                               mispredict); in this order they are right for distances 8..15 too */
                            memcpy(op, src, 8); memcpy(op + 8, src + 8, 8);
                            if (len > 16) { op += 16; src += 16; do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop); }
                        } else if (dist == 1) {
                            memset(op, *src, len);
                        } else {
                            do { *op++ = *src++; } while (op < stop);
                        }
                        op = stop;
                    } while (in_end - ip >= 16 && out_end - op >= 320);
                    /* (the entry loaded last is dropped: it consumed nothing, the careful code below looks it up again) */
                }
                REFILL();
                uint32_t e = lit_tab[BITS(LIT_TBITS)];
                if (e & E_LINK) { NEED(LIT_TBITS); DROP(LIT_TBITS); e = lit_tab[(e >> 16) + BITS((e >> 8) & 15u)]; }
                if (!(e & 0xffu)) return -1;
                NEED(e & 0xffu);
                const uint64_t saved = bitbuf;
                DROP(e & 0xffu);
                if (e & E_LITERAL) {
                    if (op >= out_end) return -1;
                    *op++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (e & E_EOB) goto block_done;
                /* the length came with its extra bits; the distance (<= 15 + 13 bits) after a fresh refill */
                const uint32_t len = XVAL(e, saved);
                REFILL();
                uint32_t d = dst_tab[BITS(DST_TBITS)];
                if (d & E_LINK) { NEED(DST_TBITS); DROP(DST_TBITS); d = dst_tab[(d >> 16) + BITS((d >> 8) & 15u)]; }
                if (!(d & 0xffu)) return -1;
                NEED(d & 0xffu);
                const uint32_t dist = XVAL(d, bitbuf);
                DROP(d & 0xffu);
                if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return -1;
                const uint8_t *src = op - dist;
                if (dist >= 8 && (size_t)(out_end - op) >= (size_t)len + 8) {
                    uint8_t *const stop = op + len;
                    do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop);   /* may overshoot by < 8 bytes inside the block */
                    op = stop;
                } else if (dist == 1) {
                    memset(op, *src, len); op += len;
                } else {
                    for (uint32_t k = 0; k < len; k++) op[k] = src[k];
                    op += len;
                }
            }
#undef XVAL
        block_done: ;
        }
        if (bfinal) break;
    }
    return op == out_end ? 0 : -1;
#undef REFILL
#undef DROP
#undef BITS
#undef NEED
}
