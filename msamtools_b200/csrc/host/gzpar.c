/* gzpar.c -- gzip writer with concurrent deflate.  See gzpar.h. */
#include "gzpar.h"
#include "crc32x.h"
#include "hostthr.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define GZP_BLOCK  ((size_t)512 << 10)     /* text bytes per deflate job */
#define GZP_DICT   32768u                  /* deflate window: what a job inherits from its predecessor */
#define GZP_STRIDE (GZP_BLOCK + GZP_BLOCK / 8 + 1024)

struct gzp {
    FILE *fp; int own_fp, threads, err;
    char *pend; size_t len, alloc, threshold;          /* pending text; a batch is deflated once it reaches `threshold` */
    unsigned char dict[GZP_DICT]; size_t dict_len;     /* the last bytes of everything already deflated */
    unsigned long crc; uint64_t total;
    unsigned char *out; size_t out_blocks;             /* GZP_STRIDE bytes per job of the current batch */
    size_t *olen; unsigned long *bcrc;
};

typedef struct { gzp *g; size_t nblk; int final, id, nthr, err; } gzp_job;

static void *gzp_worker(void *arg)
{
    gzp_job *j = arg; gzp *g = j->g;
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { j->err = 1; return NULL; }
    for (size_t k = (size_t)j->id; k < j->nblk; k += (size_t)j->nthr) {
        const size_t o = k * GZP_BLOCK, n = g->len - o < GZP_BLOCK ? g->len - o : GZP_BLOCK;
        if (deflateReset(&zs) != Z_OK) { j->err = 1; break; }
        if (k == 0) { if (g->dict_len && deflateSetDictionary(&zs, g->dict, (uInt)g->dict_len) != Z_OK) { j->err = 1; break; } }
        else if (deflateSetDictionary(&zs, (const Bytef *)g->pend + o - GZP_DICT, GZP_DICT) != Z_OK) { j->err = 1; break; }
        zs.next_in = (Bytef *)g->pend + o; zs.avail_in = (uInt)n;
        zs.next_out = g->out + k * GZP_STRIDE; zs.avail_out = (uInt)GZP_STRIDE;
        const int last = j->final && k + 1 == j->nblk;
        const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
        if ((last ? rc != Z_STREAM_END : rc != Z_OK) || zs.avail_in != 0 || zs.avail_out == 0) { j->err = 1; break; }
        g->olen[k] = GZP_STRIDE - zs.avail_out;
        g->bcrc[k] = crc32x(0, g->pend + o, n);
    }
    deflateEnd(&zs);
    return NULL;
}

static void *gzp_worker_spawned(void *arg) { worker_step_back(); return gzp_worker(arg); }

/* deflate everything pending (final: close the deflate stream, even when nothing is pending) and write it */
static int gzp_flush(gzp *g, int final)
{
    if (g->err) return -1;
    if (!g->len && !final) return 0;
    size_t nblk = (g->len + GZP_BLOCK - 1) / GZP_BLOCK;
    if (nblk == 0) nblk = 1;
    if (nblk > g->out_blocks) {
        free(g->out); free(g->olen); free(g->bcrc);
        g->out = malloc(nblk * GZP_STRIDE); g->olen = malloc(nblk * sizeof(size_t)); g->bcrc = malloc(nblk * sizeof(unsigned long));
        g->out_blocks = nblk;
        if (!g->out || !g->olen || !g->bcrc) { g->out_blocks = 0; g->err = 1; return -1; }
    }
    int t = g->threads; if ((size_t)t > nblk) t = (int)nblk;
    pthread_t th[64]; gzp_job job[64]; int spawned[64];
    for (int i = 0; i < t; i++) {
        job[i] = (gzp_job){ g, nblk, final, i, t, 0 };
        spawned[i] = i && pthread_create(&th[i], NULL, gzp_worker_spawned, &job[i]) == 0;
    }
    int err = 0;
    for (int i = 0; i < t; i++) { if (spawned[i]) pthread_join(th[i], NULL); else gzp_worker(&job[i]); err |= job[i].err; }
    if (err) { g->err = 1; return -1; }
    for (size_t k = 0; k < nblk; k++) {
        const size_t o = k * GZP_BLOCK, n = g->len - o < GZP_BLOCK ? g->len - o : GZP_BLOCK;
        if (fwrite(g->out + k * GZP_STRIDE, 1, g->olen[k], g->fp) != g->olen[k]) { g->err = 1; return -1; }
        g->crc = crc32_combine(g->crc, g->bcrc[k], (z_off_t)n);
    }
    g->total += g->len;
    if (g->len >= GZP_DICT) { memcpy(g->dict, g->pend + g->len - GZP_DICT, GZP_DICT); g->dict_len = GZP_DICT; }
    else if (g->len) {
        const size_t keep = g->dict_len + g->len > GZP_DICT ? GZP_DICT - g->len : g->dict_len;
        memmove(g->dict, g->dict + g->dict_len - keep, keep);
        memcpy(g->dict + keep, g->pend, g->len);
        g->dict_len = keep + g->len;
    }
    g->len = 0;
    return 0;
}

gzp *gzp_open(const char *path, int threads)
{
    gzp *g = calloc(1, sizeof *g);
    if (!g) return NULL;
    if (strcmp(path, "-") == 0) g->fp = stdout; else { g->fp = fopen(path, "wb"); g->own_fp = 1; }
    if (!g->fp) { free(g); return NULL; }
    g->threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
    g->threshold = (size_t)g->threads * GZP_BLOCK * 2;
    g->alloc = g->threshold + (1 << 16);
    g->pend = malloc(g->alloc);
    g->crc = crc32(0L, Z_NULL, 0);
    static const unsigned char hdr[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3 };
    if (!g->pend || fwrite(hdr, 1, 10, g->fp) != 10) { if (g->own_fp) fclose(g->fp); free(g->pend); free(g); return NULL; }
    return g;
}

char *gzp_reserve(gzp *g, size_t n)
{
    if (g->err) return NULL;
    if (g->len >= g->threshold || g->len + n > g->alloc) { if (gzp_flush(g, 0)) return NULL; }
    if (n > g->alloc) {
        char *nb = realloc(g->pend, n + (1 << 16));
        if (!nb) { g->err = 1; return NULL; }
        g->pend = nb; g->alloc = n + (1 << 16);
    }
    return g->pend + g->len;
}

int gzp_commit(gzp *g, size_t n) { g->len += n; return g->err ? -1 : 0; }

int gzp_write(gzp *g, const void *p, size_t n)
{
    const char *s = p;
    while (n) {
        size_t k = n > g->threshold ? g->threshold : n;
        char *d = gzp_reserve(g, k);
        if (!d) return -1;
        memcpy(d, s, k); g->len += k; s += k; n -= k;
    }
    return 0;
}

int gzp_puts(gzp *g, const char *s) { return gzp_write(g, s, strlen(s)); }

int gzp_printf(gzp *g, const char *fmt, ...)
{
    va_list ap;
    char *d = gzp_reserve(g, 1024);
    if (!d) return -1;
    va_start(ap, fmt);
    int n = vsnprintf(d, 1024, fmt, ap);
    va_end(ap);
    if (n < 0) { g->err = 1; return -1; }
    if (n >= 1024) {                                    /* rare: a very long name */
        d = gzp_reserve(g, (size_t)n + 1);
        if (!d) return -1;
        va_start(ap, fmt);
        vsnprintf(d, (size_t)n + 1, fmt, ap);
        va_end(ap);
    }
    g->len += (size_t)n;
    return 0;
}

int gzp_close(gzp *g)
{
    if (!g) return 0;
    int rc = gzp_flush(g, 1);
    unsigned char tr[8];
    const uint32_t c = (uint32_t)g->crc, l = (uint32_t)(g->total & 0xffffffffu);
    for (int i = 0; i < 4; i++) { tr[i] = (unsigned char)(c >> (8 * i)); tr[4 + i] = (unsigned char)(l >> (8 * i)); }
    if (!rc && fwrite(tr, 1, 8, g->fp) != 8) rc = -1;
    if (fflush(g->fp)) rc = -1;
    if (g->own_fp && fclose(g->fp)) rc = -1;
    free(g->pend); free(g->out); free(g->olen); free(g->bcrc); free(g);
    return rc;
}
