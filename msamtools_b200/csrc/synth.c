/*
 * synth.c -- synthetic name-sorted BAM record streams for tests and bench.py.
 *
 * Restates the *model* of the reference's validation generator
 * (validation/generate_synthetic_alignments.py: source genome chosen by
 * abundance x length :907-1106; records emitted "for hit in occurrences: R1, R2"
 * :982-986; source occurrence primary, others 0x100 :1004; FLAG per build_flag
 * :880-904; mismatches per mate in {0,1,2,3} w.p. {.5,.3,.1,.1} :165; tags NM, MD, AS
 * with AS = L - 2*NM :1034; MAPQ 255; QNAME sim%08d :1477; single-mate and
 * shared-locus fractions :269-286) with PE150 reads, plus soft clips and single
 * indels so that -l/-p/-z each reject a non-trivial fraction.
 * Output is the uncompressed BAM record stream + offset index that the C ABI takes.
 * Integer aux values use the smallest BAM type, as htslib does when it parses SAM text.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct synth_params {
    uint64_t seed;
    uint64_t n_records;        /* stop at the first insert boundary at or after this many records */
    uint64_t qname_base;       /* first insert number (shards use disjoint ranges)                */
    int32_t  n_refs;
    int32_t  read_len;         /* 150 */
    int32_t  insert_len;       /* 350 */
    uint32_t ref_len_min, ref_len_max;
    double   abund_sigma;      /* log-normal sigma of per-reference abundance                     */
    double   shared_fraction;  /* inserts with >1 occurrence                                      */
    double   single_fraction;  /* inserts with one mate only                                      */
    double   clip_fraction;    /* records with a soft clip                                        */
    double   indel_fraction;   /* records with one 1-3 bp I or D                                  */
    double   unmapped_fraction;/* inserts emitted as an unmapped pair (flag 4, tid -1)            */
    int32_t  max_occ;          /* cap on occurrences per shared insert (2 + Geom(0.5))            */
    int32_t  alt_noise;        /* 1: alternative occurrences may carry 0-2 extra mismatches       */
    int32_t  minimal_aux;      /* 1: no NM/MD (profile-only workloads still carry AS)             */
} synth_params;

typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t *x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static void rng_seed(rng_t *r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed); }
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_u64(rng_t *r)
{
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static double rng_unif(rng_t *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(rng_unif(r) * n); }
static double rng_normal(rng_t *r) { double u = rng_unif(r), v = rng_unif(r); if (u < 1e-300) u = 1e-300; return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

static int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

static void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

static size_t put_aux_int(uint8_t *p, const char *tag, int32_t v)
{   /* smallest type, like htslib's sam_parse1 */
    p[0] = (uint8_t)tag[0]; p[1] = (uint8_t)tag[1];
    if (v >= 0) {
        if (v <= 255) { p[2] = 'C'; p[3] = (uint8_t)v; return 4; }
        if (v <= 65535) { p[2] = 'S'; put16(p + 3, (uint32_t)v); return 5; }
        p[2] = 'I'; put32(p + 3, (uint32_t)v); return 7;
    }
    if (v >= -128) { p[2] = 'c'; p[3] = (uint8_t)(int8_t)v; return 4; }
    if (v >= -32768) { p[2] = 's'; put16(p + 3, (uint32_t)(uint16_t)(int16_t)v); return 5; }
    p[2] = 'i'; put32(p + 3, (uint32_t)v); return 7;
}

/* fills target_len[n_refs] deterministically from the seed (same for every shard) */
void synth_target_lengths(const synth_params *sp, uint32_t *target_len)
{
    rng_t r; rng_seed(&r, sp->seed ^ 0xA5A5A5A5ull);
    for (int32_t i = 0; i < sp->n_refs; i++) {
        uint32_t span = sp->ref_len_max > sp->ref_len_min ? sp->ref_len_max - sp->ref_len_min : 0;
        target_len[i] = sp->ref_len_min + (span ? rng_below(&r, span + 1) : 0);
    }
}

/* upper bound on bytes per record for buffer sizing */
size_t synth_max_record_bytes(const synth_params *sp) { return 4 + 32 + 16 + 4 * 4 + (size_t)(sp->read_len + 1) / 2 + (size_t)sp->read_len + 64; }

static const char BASES[4] = {'A', 'C', 'G', 'T'};

/* one record; returns bytes written */
static size_t emit_record(uint8_t *out, rng_t *r, const synth_params *sp, uint64_t qnum, int32_t tid, int64_t pos, uint32_t flag,
                          int32_t mate_tid, int64_t mate_pos, int32_t tlen, int nm_mis, int unmapped)
{
    const int L = sp->read_len;
    char name[32]; int lq = snprintf(name, sizeof name, "sim%08llu", (unsigned long long)qnum) + 1;
    uint32_t cigar[4]; int nc = 0;
    char md[96]; int mdl = 0;
    int clip = 0, ins = 0, del = 0, ipos = 0;
    int aligned = L;
    if (!unmapped) {
        if (rng_unif(r) < sp->clip_fraction) { clip = 1 + (int)rng_below(r, 60); if (clip > L - 30) clip = L - 30; }
        aligned = L - clip;
        if (rng_unif(r) < sp->indel_fraction) {
            int k = 1 + (int)rng_below(r, 3);
            if (rng_unif(r) < 0.5) ins = k; else del = k;
            ipos = 10 + (int)rng_below(r, (uint32_t)(aligned - 20 - k));
        }
        int clip_left = clip && rng_unif(r) < 0.5;
        if (clip && clip_left) cigar[nc++] = ((uint32_t)clip << 4) | 4;
        if (ins) { cigar[nc++] = ((uint32_t)ipos << 4) | 0; cigar[nc++] = ((uint32_t)ins << 4) | 1; cigar[nc++] = ((uint32_t)(aligned - ipos - ins) << 4) | 0; }
        else if (del) { cigar[nc++] = ((uint32_t)ipos << 4) | 0; cigar[nc++] = ((uint32_t)del << 4) | 2; cigar[nc++] = ((uint32_t)(aligned - ipos) << 4) | 0; }
        else cigar[nc++] = ((uint32_t)aligned << 4) | 0;
        if (clip && !clip_left) cigar[nc++] = ((uint32_t)clip << 4) | 4;
        /* MD over the reference-consuming M bases (aligned - ins), mismatches at distinct sorted positions */
        int mlen = aligned - ins, mp[3], nmm = nm_mis > 3 ? 3 : nm_mis;
        for (int i = 0; i < nmm; i++) {
            int ok; do { mp[i] = (int)rng_below(r, (uint32_t)mlen); ok = 1; for (int j = 0; j < i; j++) if (mp[j] == mp[i]) ok = 0;
                       } while (!ok);
        }
        for (int i = 0; i < nmm; i++) for (int j = i + 1; j < nmm; j++) if (mp[j] < mp[i]) { int t = mp[i]; mp[i] = mp[j]; mp[j] = t; }
        int mi = 0, run = 0;
        for (int cur = 0; cur < mlen; cur++) {
            if (del && cur == ipos) {                      /* "<run>^<deleted bases>" */
                mdl += snprintf(md + mdl, sizeof md - (size_t)mdl, "%d^", run); run = 0;
                for (int k = 0; k < del; k++) md[mdl++] = BASES[rng_below(r, 4)];
            }
            if (mi < nmm && mp[mi] == cur) { mdl += snprintf(md + mdl, sizeof md - (size_t)mdl, "%d%c", run, BASES[rng_below(r, 4)]); run = 0; mi++; }
            else run++;
        }
        mdl += snprintf(md + mdl, sizeof md - (size_t)mdl, "%d", run);
    }
    const int nm = nm_mis + ins + del;
    const int as = aligned - 2 * nm;
    int64_t rlen = unmapped ? 1 : (aligned - ins + del);
    uint8_t *p = out + 4;
    put32(p, (uint32_t)tid); put32(p + 4, (uint32_t)(int32_t)pos);
    p[8] = (uint8_t)lq; p[9] = unmapped ? 0 : 255;
    put16(p + 10, (uint32_t)reg2bin(pos < 0 ? 0 : pos, (pos < 0 ? 0 : pos) + (rlen > 0 ? rlen : 1)));
    put16(p + 12, (uint32_t)nc); put16(p + 14, flag);
    put32(p + 16, (uint32_t)L); put32(p + 20, (uint32_t)mate_tid); put32(p + 24, (uint32_t)(int32_t)mate_pos); put32(p + 28, (uint32_t)tlen);
    uint8_t *q = p + 32;
    memcpy(q, name, (size_t)lq); q += lq;
    for (int i = 0; i < nc; i++) { put32(q, cigar[i]); q += 4; }
    uint64_t bits = 0; int nb = 0;
    for (int i = 0; i < (L + 1) / 2; i++) {
        if (nb < 8) { bits = rng_u64(r); nb = 64; }
        uint8_t hi = (uint8_t)(1u << (bits & 3)), lo = (uint8_t)(1u << ((bits >> 2) & 3)); bits >>= 4; nb -= 4;
        q[i] = (uint8_t)(hi << 4 | ((2 * i + 1 < L) ? lo : 0));
    }
    q += (L + 1) / 2;
    memset(q, 40, (size_t)L); q += L;
    if (!unmapped) {
        if (!sp->minimal_aux) {
            q += put_aux_int(q, "NM", nm);
            q[0] = 'M'; q[1] = 'D'; q[2] = 'Z'; memcpy(q + 3, md, (size_t)mdl); q[3 + mdl] = 0; q += 4 + mdl;
        }
        q += put_aux_int(q, "AS", as);
    }
    size_t total = (size_t)(q - out);
    put32(out, (uint32_t)(total - 4));
    return total;
}

/*
 * Generate records until n_records is reached at an insert boundary, or the buffers are
 * nearly full.  Returns 0, or -1 if the buffers are too small for even one insert.
 * rec_off gets nrec+1 entries.  stats_out[0..3] = inserts, multi-occurrence inserts,
 * unmapped inserts, single-mate inserts.
 */
int synth_generate(const synth_params *sp, uint8_t *raw, size_t cap_bytes, uint64_t *rec_off, size_t cap_rec,
                   size_t *nbytes_out, size_t *nrec_out, uint64_t *stats_out)
{
    rng_t r; rng_seed(&r, sp->seed * 0x9E3779B97F4A7C15ull + sp->qname_base + 1);
    const int32_t R = sp->n_refs;
    uint32_t *tlen = malloc(sizeof(uint32_t) * (size_t)R);
    double *cum = malloc(sizeof(double) * (size_t)R);
    if (!tlen || !cum) { free(tlen); free(cum); return -2; }
    synth_target_lengths(sp, tlen);
    {   /* abundances are a property of the community (seed), not of the shard */
        rng_t ra; rng_seed(&ra, sp->seed ^ 0x5EEDull);
        double acc = 0;
        for (int32_t i = 0; i < R; i++) { double a = exp(sp->abund_sigma * rng_normal(&ra)); acc += a * tlen[i]; cum[i] = acc; }
    }
    const size_t maxrec = synth_max_record_bytes(sp);
    const int L = sp->read_len, IL = sp->insert_len;
    size_t o = 0, n = 0;
    uint64_t ins_no = 0, st[4] = {0, 0, 0, 0};
    const int max_occ = sp->max_occ > 0 ? sp->max_occ : 1;
    while (n < sp->n_records) {
        /* worst case for one insert */
        if (o + 2 * (size_t)max_occ * maxrec > cap_bytes || n + 2 * (size_t)max_occ + 1 > cap_rec) { if (n == 0) { free(tlen); free(cum); return -1; } break; }
        const uint64_t qnum = sp->qname_base + ins_no;
        ins_no++; st[0]++;
        if (rng_unif(&r) < sp->unmapped_fraction) {
            st[2]++;
            for (int mate = 1; mate <= 2; mate++) {
                uint32_t flag = 0x1 | 0x4 | 0x8 | (mate == 1 ? 0x40 : 0x80);
                rec_off[n++] = o;
                o += emit_record(raw + o, &r, sp, qnum, -1, -1, flag, -1, -1, 0, 0, 1);
            }
            continue;
        }
        int nocc = 1;
        if (rng_unif(&r) < sp->shared_fraction) { nocc = 2; while (nocc < max_occ && rng_unif(&r) < 0.5) nocc++; st[1]++; }
        int single = rng_unif(&r) < sp->single_fraction ? 1 + (int)rng_below(&r, 2) : 0;   /* 1: r1 only, 2: r2 only */
        if (single) st[3]++;
        int nm1, nm2; { double u = rng_unif(&r); nm1 = u < .5 ? 0 : u < .8 ? 1 : u < .9 ? 2 : 3; u = rng_unif(&r); nm2 = u < .5 ? 0 : u < .8 ? 1 : u < .9 ? 2 : 3; }
        for (int h = 0; h < nocc; h++) {
            /* source genome ~ abundance x length; alternatives uniform over references */
            int32_t tid;
            if (h == 0) {
                double x = rng_unif(&r) * cum[R - 1]; int32_t lo = 0, hi = R - 1;
                while (lo < hi) { int32_t mid = (lo + hi) / 2; if (cum[mid] > x) hi = mid; else lo = mid + 1; }
                tid = lo;
            } else tid = (int32_t)rng_below(&r, (uint32_t)R);
            uint32_t tl = tlen[tid];
            int64_t span = (int64_t)tl - IL; if (span < 1) span = 1;
            int64_t p1 = (int64_t)rng_below(&r, (uint32_t)span);
            int64_t p2 = p1 + IL - L; if (p2 < 0) p2 = 0;
            int extra1 = 0, extra2 = 0;
            if (h > 0 && sp->alt_noise) { extra1 = (int)rng_below(&r, 3); extra2 = (int)rng_below(&r, 3); if (rng_unif(&r) < 0.5) extra1 = extra2 = 0; }
            for (int mate = 1; mate <= 2; mate++) {
                if (single && mate != single) continue;
                int reverse = (mate == 2), mate_present = !single;
                uint32_t flag = 0x1 | (mate == 1 ? 0x40 : 0x80);
                if (reverse) flag |= 0x10;
                if (mate_present) { flag |= 0x2; if (!reverse) flag |= 0x20; } else flag |= 0x8;
                if (h > 0) flag |= 0x100;
                int nmm = (mate == 1 ? nm1 + extra1 : nm2 + extra2); if (nmm > 3) nmm = 3;
                rec_off[n++] = o;
                o += emit_record(raw + o, &r, sp, qnum, tid, mate == 1 ? p1 : p2, flag,
                                 mate_present ? tid : -1, mate_present ? (mate == 1 ? p2 : p1) : -1,
                                 mate_present ? (mate == 1 ? IL : -IL) : 0, nmm, 0);
            }
        }
    }
    rec_off[n] = o;
    *nbytes_out = o; *nrec_out = n;
    if (stats_out) memcpy(stats_out, st, sizeof st);
    free(tlen); free(cum);
    return 0;
}
