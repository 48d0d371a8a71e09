/* recwalk.c -- verified multi-threaded record index.  See recwalk.h. */
#include "recwalk.h"
#include "hostthr.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#ifdef RW_DEBUG
#include <stdio.h>
#endif

#define RW_MIN_SEGMENT ((size_t)4 << 20)    /* below this per thread the sequential walk is as fast */
#define RW_GUESS_SCAN  ((size_t)1 << 20)    /* how far into its segment a thread looks for a record start */
#define RW_GUESS_RUN   8                    /* consecutive plausible headers that make a guess */

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

/* does a BAM record header (SAM spec 4.2) plausibly start at p?  `rem` bytes are readable.  Returns its size, 0 if not */
static size_t plausible(const uint8_t *p, size_t rem, int32_t n_targets)
{
    if (rem < 36) return 0;
    const uint32_t bs = rd32(p);
    if (bs < 34 || bs > (1u << 28)) return 0;
    const int32_t tid = (int32_t)rd32(p + 4), pos = (int32_t)rd32(p + 8), lseq = (int32_t)rd32(p + 20);
    const int32_t mtid = (int32_t)rd32(p + 24), mpos = (int32_t)rd32(p + 28);
    const uint32_t lname = p[12], ncig = (uint32_t)p[16] | (uint32_t)p[17] << 8;
    if (tid < -1 || mtid < -1 || pos < -1 || mpos < -1 || lseq < 0 || lname < 1) return 0;
    if (n_targets > 0 && (tid >= n_targets || mtid >= n_targets)) return 0;
    const uint64_t need = 32ull + lname + 4ull * ncig + ((uint64_t)lseq + 1) / 2 + (uint64_t)lseq;
    if (need > bs) return 0;
    if (36 + (size_t)lname <= rem && p[36 + lname - 1] != 0) return 0;        /* read_name is NUL terminated */
    return 4 + (size_t)bs;
}

typedef struct {
    const uint8_t *raw; size_t lo, hi, len;      /* segment [lo, hi): look for a start in it, walk until >= stop */
    int32_t n_targets; int first;
    size_t start, stop_at;                       /* stop_at: the next segment's guessed start (set before the walk phase) */
    uint64_t *ends; size_t n, cap;
    size_t end_o; int status;                    /* 0 ok, 1 no guess / implausible chain, 2 oom */
} rw_seg;

static void guess_start(rw_seg *s)
{
    const size_t lim = s->lo + RW_GUESS_SCAN < s->hi ? s->lo + RW_GUESS_SCAN : s->hi;
    for (size_t p = s->lo; p < lim; p++) {
        size_t o = p; int k = 0;
        while (k < RW_GUESS_RUN) {
            if (o >= s->len) break;
            const size_t l = plausible(s->raw + o, s->len - o, s->n_targets);
            if (!l) break;
            o += l; k++;
        }
        if (k == RW_GUESS_RUN) { s->start = p; return; }
    }
    s->status = 1;
}

static void *guess_main(void *arg) { rw_seg *s = arg; if (!s->first) guess_start(s); return NULL; }

static void *walk_main(void *arg)
{
    rw_seg *s = arg;
    if (s->status) return NULL;
    const uint8_t *raw = s->raw;
    size_t o = s->start;
    s->cap = (s->stop_at - s->start) / 96 + 1024; s->n = 0;
    s->ends = malloc(s->cap * sizeof(uint64_t));
    if (!s->ends) { s->status = 2; return NULL; }
    while (o < s->stop_at && o + 4 <= s->len) {
        const uint32_t bs = rd32(raw + o);
        if (bs < 32 || bs > 0x7fffffffu) { s->status = 1; break; }           /* a true error only on the verified chain: redone sequentially */
        if (o + 4 + (size_t)bs > s->len) break;
        o += 4 + (size_t)bs;
        __builtin_prefetch(raw + o + 1024); __builtin_prefetch(raw + o + 1088);
        if (s->n == s->cap) {
            s->cap *= 2;
            uint64_t *nb = realloc(s->ends, s->cap * sizeof(uint64_t));
            if (!nb) { s->status = 2; return NULL; }
            s->ends = nb;
        }
        s->ends[s->n++] = o;
    }
    s->end_o = o;
    return NULL;
}

static void *guess_spawned(void *arg) { worker_step_back(); return guess_main(arg); }
static void *walk_spawned(void *arg) { worker_step_back(); return walk_main(arg); }

static int reserve(uint64_t **off, size_t *cap, size_t need)
{
    if (need <= *cap) return 0;
    size_t nc = *cap ? *cap : (size_t)1 << 16;
    while (nc < need) nc *= 2;
    uint64_t *nb = realloc(*off, nc * sizeof(uint64_t));
    if (!nb) return -2;
    *off = nb; *cap = nc;
    return 0;
}

static int walk_serial(const uint8_t *raw, size_t o, size_t len, uint64_t **off, size_t *n, size_t *cap)
{
    while (o + 4 <= len) {
        const uint32_t bs = rd32(raw + o);
        if (bs < 32 || bs > 0x7fffffffu) return -1;
        if (o + 4 + (size_t)bs > len) break;
        o += 4 + (size_t)bs;
        __builtin_prefetch(raw + o + 1024); __builtin_prefetch(raw + o + 1088);
        if (reserve(off, cap, *n + 2)) return -2;
        (*off)[++*n] = o;
    }
    return 0;
}

int rw_index(const uint8_t *raw, size_t from, size_t len, int32_t n_targets, int threads, uint64_t **off, size_t *n, size_t *cap)
{
    if (from >= len) return 0;
    size_t T = threads < 1 ? 1 : (threads > 64 ? 64 : (size_t)threads);
    if ((len - from) / T < RW_MIN_SEGMENT) T = (len - from) / RW_MIN_SEGMENT;
    if (T < 2) return walk_serial(raw, from, len, off, n, cap);

    rw_seg seg[64]; pthread_t th[64]; int started[64];
    const size_t span = (len - from) / T;
    for (size_t t = 0; t < T; t++) {
        memset(&seg[t], 0, sizeof seg[t]);
        seg[t].raw = raw; seg[t].len = len; seg[t].n_targets = n_targets; seg[t].first = t == 0;
        seg[t].lo = from + t * span; seg[t].hi = t + 1 == T ? len : from + (t + 1) * span;
        seg[t].start = seg[t].lo;
    }
    for (int phase = 0; phase < 2; phase++) {
        if (phase == 1)
            for (size_t t = 0; t < T; t++) {
                /* walk up to the next segment that has a guess; a segment without one is covered by its predecessor */
                size_t u = t + 1;
                while (u < T && seg[u].status) u++;
                seg[t].stop_at = u < T ? seg[u].start : len;
            }
        void *(*fn)(void *) = phase ? walk_main : guess_main;
        void *(*fn_spawned)(void *) = phase ? walk_spawned : guess_spawned;
        for (size_t t = 1; t < T; t++) started[t] = pthread_create(&th[t], NULL, fn_spawned, &seg[t]) == 0;
        fn(&seg[0]);
        for (size_t t = 1; t < T; t++) { if (started[t]) pthread_join(th[t], NULL); else fn(&seg[t]); }
    }
    /* verify and append: segment t is the true chain iff the verified chain before it arrived exactly at its start */
    int rc = 0;
    size_t o = from, t = 0;
    while (t < T) {
        if (seg[t].status == 2) { rc = -2; break; }
        if (seg[t].status || seg[t].start != o) break;                         /* no guess here, or the guess was not on the chain */
        if (reserve(off, cap, *n + seg[t].n + 2)) { rc = -2; break; }
        memcpy(*off + *n + 1, seg[t].ends, seg[t].n * sizeof(uint64_t));
        *n += seg[t].n; o = seg[t].end_o;
        size_t u = t + 1;
        while (u < T && seg[u].status == 1 && !seg[u].ends) u++;               /* segments that never walked were covered by t */
        t = u;
        if (o + 4 > len || (t < T && o > seg[t].start)) break;
        if (t < T && o < seg[t].start) break;                                  /* stopped early (partial record cannot happen mid-stream: corrupt) */
    }
    for (size_t k = 0; k < T; k++) free(seg[k].ends);
#ifdef RW_DEBUG
    fprintf(stderr, "rw_index: %zu of %zu segments verified\n", t, T);
#endif
    if (rc) return rc;
    if (t < T || o + 4 <= len) return walk_serial(raw, o, len, off, n, cap);   /* the rest (or only the tail check) sequentially */
    return 0;
}
