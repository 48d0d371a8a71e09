/* bamio.c -- SAM / BAM / BGZF I/O over zlib.  See bamio.h. */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE            /* F_SETPIPE_SZ */
#endif
#include "bamio.h"
#include "finflate.h"
#include "crc32x.h"
#include "hostthr.h"
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <zlib.h>

#define BGZF_BLOCK 0xff00          /* payload bytes per BGZF block, as htslib */
#define IN_CHUNK   (256 * 1024)

struct bio_file {
    FILE *fp; int own_fp; int writing;
    char err[256];
    /* reader */
    int compressed, z_init, z_done, in_eof, is_bam, detected;
    z_stream zs;
    uint8_t *in; size_t in_len;
    uint8_t *dec; size_t dec_len, dec_pos, dec_cap;
    char *line; size_t line_cap;
    /* parallel BGZF inflate */
    int threads, bgzf;
    uint8_t *cin; size_t cin_len;        /* view of the compressed bytes not yet consumed (inside io->mem[io_cur]) */
    struct io_ring *io; int io_cur, unbuffered;
    int block_mode;                      /* bio_set_threads was called: BGZF input is read block-wise (read-ahead thread, finflate, CRC per
                                            block) on `threads` inflate threads, 1 included; never called: zlib's streaming inflate */
    uint64_t ingest_bytes; double ingest_sec;
    uint64_t blocks_fast, blocks_zlib;   /* parallel path: blocks inflated by finflate.c / handed to zlib */
    /* name -> tid hash for SAM parsing */
    int32_t *ht; size_t ht_size; const bio_hdr *ht_hdr;
    /* writer */
    int w_bam, w_header, w_level, wz_init;
    z_stream wzs;
    uint8_t *wbuf; size_t wlen;
    uint8_t *wout; size_t *wolen;        /* bulk writer: packed blocks of one batch */
    char *fmt; size_t fmt_cap;
};

static void set_err(bio_file *f, const char *m) { snprintf(f->err, sizeof f->err, "%s", m); }
const char *bio_error(const bio_file *f) { return f ? f->err : "out of memory"; }
int bio_is_bam(const bio_file *f) { return f->writing ? f->w_bam : f->is_bam; }

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

static int grow(uint8_t **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return 0;
    size_t nc = *cap ? *cap : 4096;
    while (nc < need) nc *= 2;
    uint8_t *nb = realloc(*buf, nc);
    if (!nb) return -1;
    *buf = nb; *cap = nc;
    return 0;
}

/* ============================================================ decompressed byte stream */
static double now_sec(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

void bio_set_threads(bio_file *f, int n) { if (f) { f->threads = n < 1 ? 1 : (n > 64 ? 64 : n); f->block_mode = 1; } }
void bio_ingest_stats(const bio_file *f, uint64_t *bytes, double *seconds) { if (bytes) *bytes = f->ingest_bytes; if (seconds) *seconds = f->ingest_sec; }
void bio_inflate_stats(const bio_file *f, uint64_t *fast_blocks, uint64_t *zlib_blocks) { if (fast_blocks) *fast_blocks = f->blocks_fast; if (zlib_blocks) *zlib_blocks = f->blocks_zlib; }

/* ---- compressed input arrives through a read-ahead thread: while the workers inflate one batch, the next one is being read
 * (from a pipe that is what keeps the upstream process writing).  Two slots; the consumer's view carries the unconsumed tail
 * of a slot (less than one block) into the room reserved in front of the next one. */
#ifndef IO_BATCH                            /* (tests build with a slot of little more than one block) */
#define IO_BATCH   ((size_t)24 << 20)
#endif
#define IO_RESERVE ((size_t)1 << 17)
#define IO_MIN_POST ((size_t)8 << 20)       /* pipes: a slot holding this much is handed over the moment the writer pauses */
typedef struct io_ring {
    FILE *fp; int pipe_fd;                  /* >= 0: the input is a pipe, read(2) it directly (the stream is unbuffered, bio_open_read) */
    uint8_t *mem[2]; size_t len[2]; int full[2], eof[2];
    int quit, done, started;
    pthread_t th; pthread_mutex_t mu; pthread_cond_t cv;
} io_ring;

static void *io_main(void *arg)
{
    io_ring *r = arg;
    for (int i = 1;; i ^= 1) {
        pthread_mutex_lock(&r->mu);
        while (r->full[i] && !r->quit) pthread_cond_wait(&r->cv, &r->mu);
        const int quit = r->quit;
        pthread_mutex_unlock(&r->mu);
        if (quit) break;
        size_t got = 0; int eof = 0;
        if (r->pipe_fd >= 0) {
            /* a pipe delivers what the upstream process has written so far: waiting for a full slot would hold back, e.g., the
               header until the first records follow it.  Take what is there; stop at a full slot or when the pipe runs dry. */
            for (;;) {
                const ssize_t k = read(r->pipe_fd, r->mem[i] + IO_RESERVE + got, IO_BATCH - got);
                if (k < 0 && errno == EINTR) continue;
                if (k <= 0) { eof = 1; break; }            /* a read error ends the stream; what is missing is reported as truncation */
                got += (size_t)k;
                if (got == IO_BATCH) break;
                /* dry now: hand a big slot over at once, a small one if nothing follows within 2 ms (many tiny batches cost more) */
                struct pollfd pf = { r->pipe_fd, POLLIN, 0 };
                if (poll(&pf, 1, got >= IO_MIN_POST ? 0 : 2) <= 0) break;
            }
        } else {
            got = fread(r->mem[i] + IO_RESERVE, 1, IO_BATCH, r->fp);
            eof = got < IO_BATCH;
        }
        pthread_mutex_lock(&r->mu);
        r->len[i] = got; r->eof[i] = eof; r->full[i] = 1;
        pthread_cond_broadcast(&r->cv);
        pthread_mutex_unlock(&r->mu);
        if (eof) break;
    }
    pthread_mutex_lock(&r->mu);
    r->done = 1;
    pthread_mutex_unlock(&r->mu);
    return NULL;
}

/* the first bytes of the stream (already read for format detection) open slot 0; the thread starts unless they were all */
static int io_start(bio_file *f, const uint8_t *first, size_t n, int at_eof)
{
    io_ring *r = calloc(1, sizeof *r);
    if (!r) return -1;
    r->fp = f->fp; r->pipe_fd = -1;
    if (f->unbuffered) { struct stat sb; if (fstat(fileno(f->fp), &sb) == 0 && S_ISFIFO(sb.st_mode)) r->pipe_fd = fileno(f->fp); }
    r->mem[0] = malloc(IO_RESERVE + (n > IO_BATCH ? n : IO_BATCH));
    if (!r->mem[0]) { free(r); return -1; }
    memcpy(r->mem[0] + IO_RESERVE, first, n);
    r->len[0] = n; r->full[0] = 1; r->eof[0] = at_eof;
    pthread_mutex_init(&r->mu, NULL); pthread_cond_init(&r->cv, NULL);
    f->io = r; f->io_cur = 0; f->cin = r->mem[0] + IO_RESERVE; f->cin_len = n;
    if (!at_eof) {
        r->mem[1] = malloc(IO_RESERVE + IO_BATCH);
        if (!r->mem[1]) return -1;
        if (pthread_create(&r->th, NULL, io_main, r)) return -1;
        r->started = 1;
    }
    return 0;
}

/* append the next slot to the view (whose remaining bytes move in front of it); sets in_eof with the last one */
static void io_fetch(bio_file *f)
{
    io_ring *r = f->io;
    const int cur = f->io_cur, nxt = cur ^ 1;
    pthread_mutex_lock(&r->mu);
    while (!r->full[nxt]) pthread_cond_wait(&r->cv, &r->mu);
    pthread_mutex_unlock(&r->mu);
    uint8_t *v = r->mem[nxt] + IO_RESERVE - f->cin_len;         /* cin_len < one block <= IO_RESERVE */
    memcpy(v, f->cin, f->cin_len);
    f->cin = v; f->cin_len += r->len[nxt];
    if (r->eof[nxt]) f->in_eof = 1;
    pthread_mutex_lock(&r->mu);
    r->full[cur] = 0;
    pthread_cond_broadcast(&r->cv);
    pthread_mutex_unlock(&r->mu);
    f->io_cur = nxt;
}

static void io_stop(bio_file *f)
{
    io_ring *r = f->io;
    if (!r) return;
    if (r->started) {
        pthread_mutex_lock(&r->mu);
        r->quit = 1;
        const int done = r->done;
        pthread_cond_broadcast(&r->cv);
        pthread_mutex_unlock(&r->mu);
        /* a reader abandoned before the end of a PIPE may have its thread blocked in read(): leave it behind */
        if (!done && !f->own_fp && !f->in_eof) { pthread_detach(r->th); f->io = NULL; return; }
        pthread_join(r->th, NULL);
    }
    pthread_mutex_destroy(&r->mu); pthread_cond_destroy(&r->cv);
    free(r->mem[0]); free(r->mem[1]); free(r);
    f->io = NULL;
}

/* ---- BGZF blocks are independent gzip members: inflate a batch of them on worker threads */
typedef struct { size_t in_off, in_len, out_off; uint32_t isize, crc; } bgzf_blk;
typedef struct { const uint8_t *in; uint8_t *out; const bgzf_blk *blk; size_t nblk; int id, nthr; int err; uint64_t n_fast, n_zlib; } bgzf_job;

static void *bgzf_worker(void *arg);
static void *bgzf_worker_spawned(void *arg) { worker_step_back(); return bgzf_worker(arg); }
static void *bgzf_worker(void *arg)
{
    bgzf_job *j = arg;
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return NULL; }
    for (size_t k = (size_t)j->id; k < j->nblk; k += (size_t)j->nthr) {
        const bgzf_blk *b = &j->blk[k];
        if (b->isize == 0) continue;
        /* fast one-shot decoder first (finflate.c); anything it declines, or whose CRC does not match, goes through zlib */
        if (fi_inflate(j->in + b->in_off, b->in_len, j->out + b->out_off, b->isize) == 0 &&
            crc32x(0, j->out + b->out_off, b->isize) == b->crc) { j->n_fast++; continue; }
        j->n_zlib++;
        inflateReset(&zs);
        zs.next_in = (Bytef *)(j->in + b->in_off); zs.avail_in = (uInt)b->in_len;
        zs.next_out = j->out + b->out_off; zs.avail_out = b->isize;
        int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) { j->err = 1; break; }
        if (crc32x(0, j->out + b->out_off, b->isize) != b->crc) { j->err = 1; break; }
    }
    inflateEnd(&zs);
    return NULL;
}


/* Inflate the next batch of whole blocks.  dst == NULL: into f->dec (grown as needed), which becomes the readable window.
 * dst != NULL: straight into the caller's buffer, as many blocks as fit in dst_cap (*out_len = bytes produced; 0 with
 * return 1 means "the next block does not fit").  0 = EOF, 1 = ok, -1 = error. */
static int bgzf_batch(bio_file *f, uint8_t *dst, size_t dst_cap, size_t *out_len)
{
    for (;;) {
        /* split what the view holds into whole blocks */
        size_t p = 0, nblk = 0, out = 0, cap = 0; bgzf_blk *blk = NULL; int full = 0;
        while (p + 18 <= f->cin_len) {
            const uint8_t *h = f->cin + p;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { free(blk); set_err(f, "corrupt BGZF block header"); return -1; }
            size_t xlen = h[10] | (size_t)h[11] << 8, q = 12, bsize = 0;
            if (p + 12 + xlen > f->cin_len) break;
            while (q + 4 <= 12 + xlen) {
                size_t slen = h[q + 2] | (size_t)h[q + 3] << 8;
                if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2) bsize = (h[q + 4] | (size_t)h[q + 5] << 8) + 1;
                q += 4 + slen;
            }
            if (!bsize || bsize < 12 + xlen + 8) { free(blk); set_err(f, "corrupt BGZF block header"); return -1; }
            if (p + bsize > f->cin_len) break;
            const uint32_t isize = le32(h + bsize - 4);
            if (isize > 65536) { free(blk); set_err(f, "corrupt BGZF block header"); return -1; }      /* SAM spec 4.1: ISIZE <= 65536 */
            if (dst && out + isize > dst_cap) { full = 1; break; }
            if (nblk == cap) {
                cap = cap ? 2 * cap : 1024;
                bgzf_blk *nb = realloc(blk, cap * sizeof *blk);
                if (!nb) { free(blk); set_err(f, "out of memory"); return -1; }
                blk = nb;
            }
            blk[nblk].in_off = p + 12 + xlen; blk[nblk].in_len = bsize - 12 - xlen - 8; blk[nblk].out_off = out;
            blk[nblk].crc = le32(h + bsize - 8); blk[nblk].isize = isize;
            out += isize; nblk++; p += bsize;
        }
        if (nblk == 0) {
            free(blk);
            if (full) { *out_len = 0; return 1; }
            if (!f->in_eof) { if (f->cin_len > IO_RESERVE) { set_err(f, "corrupt BGZF block header"); return -1; } io_fetch(f); continue; }
            if (f->cin_len) { set_err(f, "truncated BGZF block"); return -1; }
            return 0;
        }
        uint8_t *target = dst;
        if (!dst) {
            if (grow(&f->dec, &f->dec_cap, out + 16)) { free(blk); set_err(f, "out of memory"); return -1; }
            target = f->dec;
        }
        int nthr = f->threads; if ((size_t)nthr > nblk) nthr = (int)nblk;
        /* a level-0 stream (what `filter -u` pipes into `profile`) is a memcpy + CRC32 per block: four threads keep up with any pipe,
           more only compete with the threads that feed it (first and last block of the batch are looked at: BTYPE 00 = stored) */
        if (nthr > 4 && blk[0].in_len && blk[nblk - 1].in_len &&
            ((f->cin[blk[0].in_off] >> 1) & 3u) == 0 && ((f->cin[blk[nblk - 1].in_off] >> 1) & 3u) == 0) nthr = 4;
        pthread_t th[64]; bgzf_job job[64]; int spawned[64];
        for (int i = 0; i < nthr; i++) {
            job[i] = (bgzf_job){ f->cin, target, blk, nblk, i, nthr, 0, 0, 0 };
            spawned[i] = i && pthread_create(&th[i], NULL, bgzf_worker_spawned, &job[i]) == 0;
        }
        int err = 0;
        for (int i = 0; i < nthr; i++) { if (spawned[i]) pthread_join(th[i], NULL); else bgzf_worker(&job[i]); err |= job[i].err; }
        for (int i = 0; i < nthr; i++) { f->blocks_fast += job[i].n_fast; f->blocks_zlib += job[i].n_zlib; }
        free(blk);
        if (err) { set_err(f, "corrupt BGZF block (inflate/CRC)"); return -1; }
        f->cin += p; f->cin_len -= p;
        if (!dst) { f->dec_pos = 0; f->dec_len = out; }
        *out_len = out;
        if (out) return 1;           /* a batch of empty (EOF-marker) blocks: look at the next one */
    }
}

static int rd_fill_bgzf(bio_file *f)
{   /* refill f->dec with the inflated payload of the next batch of whole blocks; 0 = EOF, 1 = ok, -1 = error */
    size_t out = 0;
    return bgzf_batch(f, NULL, 0, &out);
}

static int rd_fill_inner(bio_file *f);
static int rd_fill(bio_file *f)
{
    if (f->dec_pos < f->dec_len) return 1;
    const double t0 = now_sec();
    const int rc = rd_fill_inner(f);
    f->ingest_sec += now_sec() - t0;
    if (rc == 1) f->ingest_bytes += f->dec_len;
    return rc;
}

static int rd_fill_inner(bio_file *f)
{   /* make at least one more byte available in dec[dec_pos..dec_len); 0 = EOF, 1 = ok, -1 = error */
    if (f->bgzf) return rd_fill_bgzf(f);
    f->dec_pos = f->dec_len = 0;
    if (grow(&f->dec, &f->dec_cap, IN_CHUNK * 4)) { set_err(f, "out of memory"); return -1; }
    if (!f->detected) {
        f->in_len = fread(f->in, 1, IN_CHUNK, f->fp);
        if (f->in_len < IN_CHUNK) f->in_eof = 1;
        f->compressed = f->in_len >= 2 && f->in[0] == 0x1f && f->in[1] == 0x8b;
        f->detected = 1;
        if (f->compressed && f->block_mode && f->in_len >= 18 && (f->in[3] & 4) && f->in[12] == 'B' && f->in[13] == 'C') {
            f->bgzf = 1;
            if (io_start(f, f->in, f->in_len, f->in_eof)) { set_err(f, "out of memory"); return -1; }
            return rd_fill_bgzf(f);
        }
        if (f->compressed) {
            memset(&f->zs, 0, sizeof f->zs);
            if (inflateInit2(&f->zs, 15 + 32) != Z_OK) { set_err(f, "inflateInit failed"); return -1; }
            f->z_init = 1;
            f->zs.next_in = f->in; f->zs.avail_in = (uInt)f->in_len;
        } else {
            memcpy(f->dec, f->in, f->in_len); f->dec_len = f->in_len;
            return f->dec_len ? 1 : 0;
        }
    }
    if (!f->compressed) {
        f->dec_len = fread(f->dec, 1, f->dec_cap, f->fp);
        return f->dec_len ? 1 : 0;
    }
    for (;;) {
        if (f->zs.avail_in == 0) {
            if (f->in_eof) return 0;
            f->in_len = fread(f->in, 1, IN_CHUNK, f->fp);
            if (f->in_len < IN_CHUNK) f->in_eof = 1;
            if (f->in_len == 0) return 0;
            f->zs.next_in = f->in; f->zs.avail_in = (uInt)f->in_len;
        }
        f->zs.next_out = f->dec; f->zs.avail_out = (uInt)f->dec_cap;
        int rc = inflate(&f->zs, Z_NO_FLUSH);
        if (rc == Z_STREAM_END) {
            /* BGZF is a series of gzip members: restart on the next one */
            if (inflateReset(&f->zs) != Z_OK) { set_err(f, "inflateReset failed"); return -1; }
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) { set_err(f, "corrupt gzip/BGZF stream"); return -1; }
        f->dec_len = f->dec_cap - f->zs.avail_out;
        if (f->dec_len) return 1;
        if (rc == Z_BUF_ERROR && f->zs.avail_in == 0 && f->in_eof) return 0;
    }
}

static int rd_read(bio_file *f, uint8_t *dst, size_t n)
{   /* 1 = got n bytes, 0 = clean EOF before the first byte, -1 = error / truncated */
    size_t got = 0;
    while (got < n) {
        int rc = rd_fill(f);
        if (rc < 0) return -1;
        if (rc == 0) { if (got == 0) return 0; set_err(f, "truncated file"); return -1; }
        size_t k = f->dec_len - f->dec_pos; if (k > n - got) k = n - got;
        memcpy(dst + got, f->dec + f->dec_pos, k);
        f->dec_pos += k; got += k;
    }
    return 1;
}

int bio_read_raw(bio_file *f, uint8_t *buf, size_t cap, size_t *len)
{   /* append decompressed stream bytes at buf + *len (see bamio.h); 1 = appended something, 0 = EOF, -1 = error, 2 = buffer full */
    if (*len >= cap) return 2;
    if (f->dec_pos < f->dec_len) {                         /* what the record-wise reader left in the window */
        size_t k = f->dec_len - f->dec_pos; if (k > cap - *len) k = cap - *len;
        memcpy(buf + *len, f->dec + f->dec_pos, k);
        f->dec_pos += k; *len += k;
        return 1;
    }
    if (f->bgzf) {                                         /* parallel BGZF: inflate whole blocks straight into the caller's buffer */
        const double t0 = now_sec();
        size_t out = 0;
        const int rc = bgzf_batch(f, buf + *len, cap - *len, &out);
        f->ingest_sec += now_sec() - t0;
        if (rc <= 0) return rc;
        if (out == 0) return 2;
        f->ingest_bytes += out; *len += out;
        return 1;
    }
    const int rc = rd_fill(f);                             /* streaming inflate / plain file: through the window */
    if (rc <= 0) return rc;
    size_t k = f->dec_len - f->dec_pos; if (k > cap - *len) k = cap - *len;
    memcpy(buf + *len, f->dec + f->dec_pos, k);
    f->dec_pos += k; *len += k;
    return 1;
}

static int rd_peek4(bio_file *f, uint8_t out[4])
{   /* only used right at the start of the stream, where 4 bytes are contiguous in dec */
    int rc = rd_fill(f);
    if (rc <= 0) return rc;
    if (f->dec_len - f->dec_pos < 4) return 0;
    memcpy(out, f->dec + f->dec_pos, 4);
    return 1;
}

static long rd_line(bio_file *f)
{   /* next text line into f->line (NUL terminated, no newline); returns length, -1 EOF, -2 error */
    size_t l = 0;
    for (;;) {
        int rc = rd_fill(f);
        if (rc < 0) return -2;
        if (rc == 0) { if (l == 0) return -1; break; }
        uint8_t *s = f->dec + f->dec_pos; size_t avail = f->dec_len - f->dec_pos;
        uint8_t *nl = memchr(s, '\n', avail);
        size_t k = nl ? (size_t)(nl - s) : avail;
        if (l + k + 1 > f->line_cap) {
            size_t nc = f->line_cap ? f->line_cap : 1024; while (nc < l + k + 1) nc *= 2;
            char *nb = realloc(f->line, nc); if (!nb) { set_err(f, "out of memory"); return -2; }
            f->line = nb; f->line_cap = nc;
        }
        memcpy(f->line + l, s, k); l += k;
        f->dec_pos += k + (nl ? 1 : 0);
        if (nl) break;
    }
    if (l && f->line[l - 1] == '\r') l--;
    if (!f->line) { f->line = malloc(16); f->line_cap = 16; }
    f->line[l] = 0;
    return (long)l;
}

/* a pipe between two msamtools processes carries gigabytes: ask for a 1 MB pipe buffer instead of 64 KB (fewer wake-ups per
 * byte; silently keeps the default where the kernel says no or the descriptor is not a pipe) */
static void widen_pipe(FILE *fp)
{
#ifdef F_SETPIPE_SZ
    if (fp) (void)fcntl(fileno(fp), F_SETPIPE_SZ, 1 << 20);
#else
    (void)fp;
#endif
}

bio_file *bio_open_read(const char *path)
{
    bio_file *f = calloc(1, sizeof *f);
    if (!f) return NULL;
    if (strcmp(path, "-") == 0) {
        f->fp = stdin; widen_pipe(stdin);
        f->unbuffered = setvbuf(stdin, NULL, _IONBF, 0) == 0;      /* every read asks for >= 64 KB; lets the read-ahead thread use read(2) on a pipe */
    } else { f->fp = fopen(path, "rb"); f->own_fp = 1; }
    if (!f->fp) { free(f); return NULL; }
    f->in = malloc(IN_CHUNK);
    if (!f->in) { if (f->own_fp) fclose(f->fp); free(f); return NULL; }
    f->threads = 1;
    return f;
}

/* ============================================================ header */
static char *xstrndup(const char *s, size_t n) { char *r = malloc(n + 1); if (r) { memcpy(r, s, n); r[n] = 0; } return r; }

bio_hdr *bio_hdr_parse_text(const char *text, size_t l_text)
{
    bio_hdr *h = calloc(1, sizeof *h);
    if (!h) return NULL;
    h->text = xstrndup(text, l_text); h->l_text = l_text;
    size_t cap = 0;
    const char *p = text, *end = text + l_text;
    while (p < end) {
        const char *nl = memchr(p, '\n', (size_t)(end - p)); if (!nl) nl = end;
        if (nl - p >= 3 && p[0] == '@' && p[1] == 'S' && p[2] == 'Q') {
            const char *sn = NULL, *q = p; size_t lsn = 0; long ln = -1;
            while (q < nl) {
                const char *tab = memchr(q, '\t', (size_t)(nl - q)); if (!tab) tab = nl;
                if (tab - q > 3 && q[2] == ':') {
                    if (q[0] == 'S' && q[1] == 'N') { sn = q + 3; lsn = (size_t)(tab - q - 3); }
                    if (q[0] == 'L' && q[1] == 'N') ln = strtol(q + 3, NULL, 10);
                }
                q = tab + 1;
            }
            if (sn && ln >= 0) {
                if ((size_t)h->n_targets == cap) {
                    cap = cap ? cap * 2 : 64;
                    h->target_name = realloc(h->target_name, cap * sizeof(char *));
                    h->target_len = realloc(h->target_len, cap * sizeof(uint32_t));
                }
                h->target_name[h->n_targets] = xstrndup(sn, lsn);
                h->target_len[h->n_targets] = (uint32_t)ln;
                h->n_targets++;
            }
        }
        p = nl + 1;
    }
    return h;
}

bio_hdr *bio_hdr_dup(const bio_hdr *s)
{
    bio_hdr *h = calloc(1, sizeof *h);
    if (!h) return NULL;
    h->text = xstrndup(s->text ? s->text : "", s->l_text); h->l_text = s->l_text; h->n_targets = s->n_targets;
    h->target_name = malloc(sizeof(char *) * (size_t)(s->n_targets ? s->n_targets : 1));
    h->target_len = malloc(sizeof(uint32_t) * (size_t)(s->n_targets ? s->n_targets : 1));
    size_t total = 1;
    for (int32_t i = 0; i < s->n_targets; i++) total += strlen(s->target_name[i]) + 1;
    h->name_arena = malloc(total);
    if (!h->text || !h->target_name || !h->target_len || !h->name_arena) { h->n_targets = 0; bio_hdr_free(h); return NULL; }
    char *a = h->name_arena;
    for (int32_t i = 0; i < s->n_targets; i++) {
        const size_t l = strlen(s->target_name[i]) + 1;
        memcpy(a, s->target_name[i], l); h->target_name[i] = a; a += l;
        h->target_len[i] = s->target_len[i];
    }
    return h;
}

void bio_hdr_free(bio_hdr *h)
{
    if (!h) return;
    if (h->name_arena) free(h->name_arena);
    else for (int32_t i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h);
}

char *bio_hdr_find_hd_tag(const bio_hdr *h, const char *tag)
{
    const char *p = h->text, *end = h->text + h->l_text;
    while (p && p < end) {
        const char *nl = memchr(p, '\n', (size_t)(end - p)); if (!nl) nl = end;
        if (nl - p >= 3 && p[0] == '@' && p[1] == 'H' && p[2] == 'D') {
            const char *q = p;
            while (q < nl) {
                const char *tab = memchr(q, '\t', (size_t)(nl - q)); if (!tab) tab = nl;
                if (tab - q >= 3 && q[0] == tag[0] && q[1] == tag[1] && q[2] == ':') return xstrndup(q + 3, (size_t)(tab - q - 3));
                q = tab + 1;
            }
            return NULL;
        }
        p = nl + 1;
    }
    return NULL;
}

int bio_hdr_add_pg(bio_hdr *h, const char *name, const char *pn, const char *vn, const char *cl, const char *ds)
{
    /* collect @PG IDs and the IDs referenced by PP */
    enum { MAXPG = 256 };
    char *ids[MAXPG], *pps[MAXPG]; int nid = 0, npp = 0;
    const char *p = h->text, *end = h->text + h->l_text;
    while (p < end) {
        const char *nl = memchr(p, '\n', (size_t)(end - p)); if (!nl) nl = end;
        if (nl - p >= 3 && p[0] == '@' && p[1] == 'P' && p[2] == 'G') {
            const char *q = p;
            while (q < nl) {
                const char *tab = memchr(q, '\t', (size_t)(nl - q)); if (!tab) tab = nl;
                if (tab - q >= 3 && q[2] == ':') {
                    if (q[0] == 'I' && q[1] == 'D' && nid < MAXPG) ids[nid++] = xstrndup(q + 3, (size_t)(tab - q - 3));
                    if (q[0] == 'P' && q[1] == 'P' && npp < MAXPG) pps[npp++] = xstrndup(q + 3, (size_t)(tab - q - 3));
                }
                q = tab + 1;
            }
        }
        p = nl + 1;
    }
    /* chain tails = IDs nobody points at; one new line per tail (htslib), or one line without PP */
    char *tails[MAXPG]; int nt = 0;
    for (int i = 0; i < nid; i++) { int ref = 0; for (int j = 0; j < npp; j++) if (!strcmp(ids[i], pps[j])) ref = 1; if (!ref) tails[nt++] = ids[i]; }
    int nadd = nt ? nt : 1, rc = 0;
    for (int a = 0; a < nadd && !rc; a++) {
        char id[300]; int suffix = 0;
        for (;;) {          /* unique ID: name, name.1, name.2, ... */
            if (suffix) snprintf(id, sizeof id, "%s.%d", name, suffix); else snprintf(id, sizeof id, "%s", name);
            int clash = 0; for (int i = 0; i < nid; i++) if (!strcmp(ids[i], id)) clash = 1;
            if (!clash) break;
            suffix++;
        }
        size_t need = strlen(id) + strlen(pn) + strlen(vn) + strlen(cl) + strlen(ds) + 64 + (nt ? strlen(tails[a]) : 0);
        char *line = malloc(need);
        if (!line) { rc = -1; break; }
        int n = snprintf(line, need, "@PG\tID:%s\tPN:%s", id, pn);
        if (nt) n += snprintf(line + n, need - (size_t)n, "\tPP:%s", tails[a]);
        n += snprintf(line + n, need - (size_t)n, "\tVN:%s\tCL:%s\tDS:%s\n", vn, cl, ds);
        int need_nl = h->l_text && h->text[h->l_text - 1] != '\n';
        char *nt_ = realloc(h->text, h->l_text + (size_t)n + 2);
        if (!nt_) { free(line); rc = -1; break; }
        h->text = nt_;
        if (need_nl) h->text[h->l_text++] = '\n';
        memcpy(h->text + h->l_text, line, (size_t)n); h->l_text += (size_t)n; h->text[h->l_text] = 0;
        if (nid < MAXPG) ids[nid++] = xstrndup(id, strlen(id));
        free(line);
    }
    for (int i = 0; i < nid; i++) free(ids[i]);
    for (int i = 0; i < npp; i++) free(pps[i]);
    return rc;
}

/* name -> tid: open addressing over FNV-1a */
static uint64_t fnv(const char *s, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 1099511628211ull; } return h; }
static void build_ht(bio_file *f, const bio_hdr *h)
{
    free(f->ht);
    size_t sz = 16; while (sz < (size_t)h->n_targets * 2 + 1) sz *= 2;
    f->ht = malloc(sz * sizeof(int32_t)); f->ht_size = sz; f->ht_hdr = h;
    for (size_t i = 0; i < sz; i++) f->ht[i] = -1;
    for (int32_t t = 0; t < h->n_targets; t++) {
        size_t k = fnv(h->target_name[t], strlen(h->target_name[t])) & (sz - 1);
        while (f->ht[k] >= 0) { if (!strcmp(h->target_name[f->ht[k]], h->target_name[t])) break; k = (k + 1) & (sz - 1); }
        if (f->ht[k] < 0) f->ht[k] = t;
    }
}
static int32_t lookup_tid(bio_file *f, const bio_hdr *h, const char *s, size_t n)
{
    if (!f) { for (int32_t t = 0; t < h->n_targets; t++) if (strlen(h->target_name[t]) == n && !memcmp(h->target_name[t], s, n)) return t; return -2; }
    if (f->ht_hdr != h) build_ht(f, h);
    size_t k = fnv(s, n) & (f->ht_size - 1);
    while (f->ht[k] >= 0) {
        const char *nm = h->target_name[f->ht[k]];
        if (strlen(nm) == n && !memcmp(nm, s, n)) return f->ht[k];
        k = (k + 1) & (f->ht_size - 1);
    }
    return -2;
}
int bio_hdr_tid(const bio_hdr *h, const char *name) { return lookup_tid(NULL, h, name, strlen(name)); }

bio_hdr *bio_read_header(bio_file *f)
{
    uint8_t magic[4];
    int rc = rd_peek4(f, magic);
    if (rc < 0) return NULL;
    if (rc == 1 && !memcmp(magic, "BAM\1", 4)) {
        uint8_t b[8];
        f->is_bam = 1;
        if (rd_read(f, b, 8) != 1) return NULL;
        uint32_t l_text = le32(b + 4);
        if (l_text > 0x7fffffffu) { set_err(f, "corrupt BAM header (l_text)"); return NULL; }
        char *text = malloc((size_t)l_text + 1);
        if (!text || (l_text && rd_read(f, (uint8_t *)text, l_text) != 1)) { free(text); set_err(f, text ? "truncated BAM header" : "out of memory"); return NULL; }
        text[l_text] = 0;
        size_t tl = strlen(text);
        if (rd_read(f, b, 4) != 1) { free(text); set_err(f, "truncated BAM header"); return NULL; }
        int32_t n_ref = (int32_t)le32(b);
        if (n_ref < 0) { free(text); set_err(f, "corrupt BAM header (n_ref)"); return NULL; }
        bio_hdr *h = calloc(1, sizeof *h);
        if (!h) { free(text); set_err(f, "out of memory"); return NULL; }
        h->text = text; h->l_text = tl; h->n_targets = 0;
        h->target_name = calloc((size_t)(n_ref ? n_ref : 1), sizeof(char *));
        h->target_len = calloc((size_t)(n_ref ? n_ref : 1), sizeof(uint32_t));
        if (!h->target_name || !h->target_len) { bio_hdr_free(h); set_err(f, "out of memory"); return NULL; }
        /* names go into one arena (grown geometrically; pointers are fixed up at the end), read straight out of the
           decompressed window whenever a whole entry is inside it */
        size_t acap = (size_t)(n_ref < (1 << 20) ? n_ref : (1 << 20)) * 16 + 64, alen = 0;      /* n_ref is only a claim until the entries have been read */
        char *arena = malloc(acap);
        size_t *aoff = malloc(sizeof(size_t) * (size_t)(n_ref ? n_ref : 1));
        if (!arena || !aoff) { free(arena); free(aoff); bio_hdr_free(h); set_err(f, "out of memory"); return NULL; }
        for (int32_t i = 0; i < n_ref; i++) {
            uint32_t ln; const uint8_t *src = NULL;
            if (f->dec_len - f->dec_pos >= 4 && (ln = le32(f->dec + f->dec_pos)) <= (1u << 20) && f->dec_len - f->dec_pos >= 8 + (size_t)ln) {
                src = f->dec + f->dec_pos + 4;                              /* whole entry in the window */
            } else {
                if (rd_read(f, b, 4) != 1) { free(arena); free(aoff); bio_hdr_free(h); set_err(f, "truncated BAM header"); return NULL; }
                ln = le32(b);
            }
            if (ln == 0 || ln > (1u << 20)) { free(arena); free(aoff); bio_hdr_free(h); set_err(f, "corrupt BAM header (l_name)"); return NULL; }
            if (alen + ln + 1 > acap) {
                while (alen + ln + 1 > acap) acap *= 2;
                char *na = realloc(arena, acap);
                if (!na) { free(arena); free(aoff); bio_hdr_free(h); set_err(f, "out of memory"); return NULL; }
                arena = na;
            }
            if (src) { memcpy(arena + alen, src, ln); h->target_len[i] = le32(src + ln); f->dec_pos += 8 + (size_t)ln; }
            else {
                if (rd_read(f, (uint8_t *)arena + alen, ln) != 1 || rd_read(f, b, 4) != 1) { free(arena); free(aoff); bio_hdr_free(h); set_err(f, "truncated BAM header"); return NULL; }
                h->target_len[i] = le32(b);
            }
            arena[alen + ln] = 0;
            aoff[i] = alen; alen += (size_t)ln + 1;
        }
        for (int32_t i = 0; i < n_ref; i++) h->target_name[i] = arena + aoff[i];
        h->name_arena = arena; h->n_targets = n_ref;
        free(aoff);
        return h;
    }
    /* SAM text: header lines start with '@'; the first alignment line is kept for the first bio_read_record */
    size_t cap = 4096, l = 0; char *text = malloc(cap);
    for (;;) {
        int frc = rd_fill(f);
        if (frc < 0) { free(text); return NULL; }
        if (frc == 0 || f->dec[f->dec_pos] != '@') break;
        long n = rd_line(f);
        if (n < 0) break;
        if (l + (size_t)n + 2 > cap) { while (l + (size_t)n + 2 > cap) cap *= 2; text = realloc(text, cap); }
        memcpy(text + l, f->line, (size_t)n); l += (size_t)n; text[l++] = '\n';
    }
    text[l] = 0;
    bio_hdr *h = bio_hdr_parse_text(text, l);
    free(text);
    return h;
}

/* ============================================================ SAM text <-> BAM record */
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

static const char SEQ_CODES[] = "=ACMGRSVTWYHKDBN";
static int seq_code(int c)
{
    c = toupper(c);
    const char *p = strchr(SEQ_CODES, c);
    return (p && c) ? (int)(p - SEQ_CODES) : 15;
}

#define PERR(msg) do { if (err) snprintf(err, lerr, "%s", msg); return -1; } while (0)

static int parse_sam_line(bio_file *f, const char *line, size_t l, const bio_hdr *h, uint8_t **buf, size_t *cap, size_t *len, char *err, size_t lerr)
{
    const char *fld[11]; size_t fl[11];
    const char *p = line, *end = line + l;
    for (int i = 0; i < 11; i++) {
        if (p > end) PERR("SAM record has fewer than 11 fields");
        const char *tab = memchr(p, '\t', (size_t)(end - p));
        if (!tab) { if (i < 10) PERR("SAM record has fewer than 11 fields"); tab = end; }
        fld[i] = p; fl[i] = (size_t)(tab - p); p = tab + 1;
    }
    const char *aux = p <= end ? p : end;      /* start of optional fields (may be == end) */
    size_t lq = fl[0] + 1;
    if (lq > 255) PERR("query name too long");
    uint32_t flag = (uint32_t)strtoul(fld[1], NULL, 0);
    int32_t tid = (fl[2] == 1 && fld[2][0] == '*') ? -1 : lookup_tid(f, h, fld[2], fl[2]);
    if (tid == -2) PERR("unrecognised reference name in RNAME");
    int64_t pos = strtoll(fld[3], NULL, 10) - 1;
    uint32_t mapq = (uint32_t)strtoul(fld[4], NULL, 10);
    /* cigar */
    size_t ncig = 0;
    if (!(fl[5] == 1 && fld[5][0] == '*')) for (size_t i = 0; i < fl[5]; i++) if (!isdigit((unsigned char)fld[5][i])) ncig++;
    if (ncig > 65535) PERR("CIGAR with more than 65535 operations is not supported");
    int32_t mtid = (fl[6] == 1 && fld[6][0] == '*') ? -1 : (fl[6] == 1 && fld[6][0] == '=') ? tid : lookup_tid(f, h, fld[6], fl[6]);
    if (mtid == -2) PERR("unrecognised reference name in RNEXT");
    int64_t mpos = strtoll(fld[7], NULL, 10) - 1;
    int64_t isize = strtoll(fld[8], NULL, 10);
    size_t lseq = (fl[9] == 1 && fld[9][0] == '*') ? 0 : fl[9];
    if (!(fl[10] == 1 && fld[10][0] == '*') && fl[10] != lseq) PERR("SEQ and QUAL are of different length");
    size_t auxmax = (size_t)(end - aux) + 16;
    size_t need = *len + 36 + lq + 4 * ncig + (lseq + 1) / 2 + lseq + auxmax * 2;
    if (grow(buf, cap, need)) PERR("out of memory");
    uint8_t *r = *buf + *len, *q = r + 36;
    memcpy(q, fld[0], fl[0]); q[fl[0]] = 0; q += lq;
    int64_t rlen = 0;
    {
        const char *c = fld[5], *ce = fld[5] + fl[5];
        for (size_t k = 0; k < ncig; k++) {
            char *ep; unsigned long n = strtoul(c, &ep, 10);
            if (ep >= ce) PERR("malformed CIGAR");
            const char *o = strchr("MIDNSHP=XB", *ep);
            if (!o || !*ep) PERR("unrecognised CIGAR operator");
            int op = (int)(o - "MIDNSHP=XB");
            put32(q, (uint32_t)(n << 4) | (uint32_t)op); q += 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += (int64_t)n;
            c = ep + 1;
        }
    }
    memset(q, 0, (lseq + 1) / 2);
    for (size_t i = 0; i < lseq; i++) q[i >> 1] |= (uint8_t)(seq_code((unsigned char)fld[9][i]) << ((~i & 1) << 2));
    q += (lseq + 1) / 2;
    if (fl[10] == 1 && fld[10][0] == '*') memset(q, 0xff, lseq);
    else for (size_t i = 0; i < lseq; i++) q[i] = (uint8_t)(fld[10][i] - 33);
    q += lseq;
    /* optional fields */
    p = aux;
    while (p < end) {
        const char *tab = memchr(p, '\t', (size_t)(end - p)); if (!tab) tab = end;
        size_t n = (size_t)(tab - p);
        if (n == 0) { p = tab + 1; continue; }
        if (n < 5 || p[2] != ':' || p[4] != ':') PERR("malformed optional field");
        q[0] = (uint8_t)p[0]; q[1] = (uint8_t)p[1];
        char ty = p[3]; const char *v = p + 5; size_t vl = n - 5;
        if (ty == 'A') { q[2] = 'A'; q[3] = (uint8_t)v[0]; q += 4; }
        else if (ty == 'i' || ty == 'I') {
            long long x = strtoll(v, NULL, 10);
            if (x < 0) {
                if (x >= -128) { q[2] = 'c'; q[3] = (uint8_t)(int8_t)x; q += 4; }
                else if (x >= -32768) { q[2] = 's'; put16(q + 3, (uint32_t)(uint16_t)(int16_t)x); q += 5; }
                else { q[2] = 'i'; put32(q + 3, (uint32_t)(int32_t)x); q += 7; }
            } else {
                if (x <= 255) { q[2] = 'C'; q[3] = (uint8_t)x; q += 4; }
                else if (x <= 65535) { q[2] = 'S'; put16(q + 3, (uint32_t)x); q += 5; }
                else { q[2] = 'I'; put32(q + 3, (uint32_t)x); q += 7; }
            }
        } else if (ty == 'f') { float x = strtof(v, NULL); q[2] = 'f'; memcpy(q + 3, &x, 4); q += 7; }
        else if (ty == 'd') { double x = strtod(v, NULL); q[2] = 'd'; memcpy(q + 3, &x, 8); q += 11; }
        else if (ty == 'Z' || ty == 'H') { q[2] = (uint8_t)ty; memcpy(q + 3, v, vl); q[3 + vl] = 0; q += 4 + vl; }
        else if (ty == 'B') {
            if (vl < 1) PERR("malformed B field");
            char st = v[0]; size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
            if (!es) PERR("unrecognised B subtype");
            size_t cnt = 0; for (size_t i = 1; i < vl; i++) if (v[i] == ',') cnt++;
            size_t used = (size_t)(q - *buf);
            if (grow(buf, cap, used + 8 + cnt * es + auxmax * 2)) PERR("out of memory");
            r = *buf + *len; q = *buf + used;
            q[2] = 'B'; q[3] = (uint8_t)st; put32(q + 4, (uint32_t)cnt); q += 8;
            const char *c = v + 1;
            for (size_t i = 0; i < cnt; i++) {
                c++; char *ep;
                if (st == 'f') { float x = strtof(c, &ep); memcpy(q, &x, 4); }
                else { long long x = strtoll(c, &ep, 10); if (es == 1) q[0] = (uint8_t)x; else if (es == 2) put16(q, (uint32_t)(uint16_t)x); else put32(q, (uint32_t)x); }
                q += es; c = ep;
            }
        } else PERR("unrecognised optional field type");
        p = tab + 1;
    }
    int64_t endpos = pos + ((flag & 4) || rlen == 0 ? 1 : rlen);
    size_t total = (size_t)(q - r);
    put32(r, (uint32_t)(total - 4));
    put32(r + 4, (uint32_t)tid); put32(r + 8, (uint32_t)(int32_t)pos);
    r[12] = (uint8_t)lq; r[13] = (uint8_t)mapq; put16(r + 14, (uint32_t)reg2bin(pos, endpos));
    put16(r + 16, (uint32_t)ncig); put16(r + 18, flag);
    put32(r + 20, (uint32_t)lseq); put32(r + 24, (uint32_t)mtid); put32(r + 28, (uint32_t)(int32_t)mpos); put32(r + 32, (uint32_t)(int32_t)isize);
    *len += total;
    return 0;
}

int bio_sam_parse(const char *line, size_t l, const bio_hdr *h, uint8_t **buf, size_t *cap, size_t *len, char *err, size_t lerr)
{
    return parse_sam_line(NULL, line, l, h, buf, cap, len, err, lerr);
}

int bio_read_record(bio_file *f, const bio_hdr *h, uint8_t **buf, size_t *cap, size_t *len)
{
    if (f->is_bam) {
        uint8_t b[4];
        int rc = rd_read(f, b, 4);
        if (rc <= 0) return rc;
        uint32_t bs = le32(b);
        if (bs < 32 || bs > 0x7fffffffu) { set_err(f, "corrupt BAM record"); return -1; }
        if (grow(buf, cap, *len + 4 + bs)) { set_err(f, "out of memory"); return -1; }
        memcpy(*buf + *len, b, 4);
        if (rd_read(f, *buf + *len + 4, bs) != 1) { set_err(f, "truncated BAM record"); return -1; }
        *len += 4 + (size_t)bs;
        return 1;
    }
    for (;;) {
        long n = rd_line(f);
        if (n == -1) return 0;
        if (n < 0) return -1;
        if (n == 0 || f->line[0] == '@') continue;
        if (parse_sam_line(f, f->line, (size_t)n, h, buf, cap, len, f->err, sizeof f->err)) return -1;
        return 1;
    }
}

static int sputs(char **out, size_t *cap, size_t *l, const char *s, size_t n)
{
    if (*l + n + 1 > *cap) { size_t nc = *cap ? *cap : 1024; while (nc < *l + n + 1) nc *= 2; char *nb = realloc(*out, nc); if (!nb) return -1; *out = nb; *cap = nc; }
    memcpy(*out + *l, s, n); *l += n; (*out)[*l] = 0;
    return 0;
}
static int sprintf_i(char **out, size_t *cap, size_t *l, long long v) { char b[32]; int n = snprintf(b, sizeof b, "%lld", v); return sputs(out, cap, l, b, (size_t)n); }
static int sprintf_g(char **out, size_t *cap, size_t *l, double v) { char b[64]; int n = snprintf(b, sizeof b, "%g", v); return sputs(out, cap, l, b, (size_t)n); }

int bio_sam_format(const uint8_t *r, size_t len, const bio_hdr *h, char **out, size_t *cap, size_t *l)
{
    if (len < 36) return -1;
    int32_t tid = (int32_t)le32(r + 4), pos = (int32_t)le32(r + 8), lseq = (int32_t)le32(r + 20);
    uint32_t lq = r[12], mapq = r[13], nc = le16(r + 16), flag = le16(r + 18);
    int32_t mtid = (int32_t)le32(r + 24), mpos = (int32_t)le32(r + 28), isize = (int32_t)le32(r + 32);
    const uint8_t *q = r + 36, *end = r + len;
    if (36 + (size_t)lq + 4 * (size_t)nc + ((size_t)lseq + 1) / 2 + (size_t)lseq > len) return -1;
    sputs(out, cap, l, (const char *)q, lq ? lq - 1 : 0); q += lq;
    sputs(out, cap, l, "\t", 1); sprintf_i(out, cap, l, flag); sputs(out, cap, l, "\t", 1);
    if (tid >= 0 && tid < h->n_targets) sputs(out, cap, l, h->target_name[tid], strlen(h->target_name[tid])); else sputs(out, cap, l, "*", 1);
    sputs(out, cap, l, "\t", 1); sprintf_i(out, cap, l, (long long)pos + 1);
    sputs(out, cap, l, "\t", 1); sprintf_i(out, cap, l, mapq); sputs(out, cap, l, "\t", 1);
    if (nc == 0) sputs(out, cap, l, "*", 1);
    for (uint32_t k = 0; k < nc; k++) { uint32_t c = le32(q + 4 * k); sprintf_i(out, cap, l, c >> 4); char o = "MIDNSHP=XB??????"[c & 15]; sputs(out, cap, l, &o, 1); }
    q += 4 * nc;
    sputs(out, cap, l, "\t", 1);
    if (mtid < 0) sputs(out, cap, l, "*", 1);
    else if (mtid == tid) sputs(out, cap, l, "=", 1);
    else if (mtid < h->n_targets) sputs(out, cap, l, h->target_name[mtid], strlen(h->target_name[mtid])); else sputs(out, cap, l, "*", 1);
    sputs(out, cap, l, "\t", 1); sprintf_i(out, cap, l, (long long)mpos + 1);
    sputs(out, cap, l, "\t", 1); sprintf_i(out, cap, l, isize); sputs(out, cap, l, "\t", 1);
    if (lseq == 0) sputs(out, cap, l, "*\t*", 3);
    else {
        if (*l + 2 * (size_t)lseq + 4 > *cap) { size_t nc2 = *cap ? *cap : 1024; while (nc2 < *l + 2 * (size_t)lseq + 4) nc2 *= 2; char *nb = realloc(*out, nc2); if (!nb) return -1; *out = nb; *cap = nc2; }
        char *o = *out + *l;
        for (int32_t i = 0; i < lseq; i++) o[i] = SEQ_CODES[(q[i >> 1] >> ((~i & 1) << 2)) & 15];
        o += lseq; *o++ = '\t'; q += ((size_t)lseq + 1) / 2;
        if (q[0] == 0xff) { *o++ = '*'; } else for (int32_t i = 0; i < lseq; i++) *o++ = (char)(q[i] + 33);
        *l = (size_t)(o - *out); (*out)[*l] = 0;
        q += lseq;
    }
    while (end - q >= 3) {
        char hd[8]; int n = snprintf(hd, sizeof hd, "\t%c%c:", q[0], q[1]); sputs(out, cap, l, hd, (size_t)n);
        uint8_t ty = q[2]; q += 3;
        if (ty == 'A') { char b[3] = {'A', ':', (char)q[0]}; sputs(out, cap, l, b, 3); q += 1; }
        else if (ty == 'c') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, (int8_t)q[0]); q += 1; }
        else if (ty == 'C') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, q[0]); q += 1; }
        else if (ty == 's') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, (int16_t)le16(q)); q += 2; }
        else if (ty == 'S') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, le16(q)); q += 2; }
        else if (ty == 'i') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, (int32_t)le32(q)); q += 4; }
        else if (ty == 'I') { sputs(out, cap, l, "i:", 2); sprintf_i(out, cap, l, le32(q)); q += 4; }
        else if (ty == 'f') { float x; memcpy(&x, q, 4); sputs(out, cap, l, "f:", 2); sprintf_g(out, cap, l, x); q += 4; }
        else if (ty == 'd') { double x; memcpy(&x, q, 8); sputs(out, cap, l, "d:", 2); sprintf_g(out, cap, l, x); q += 8; }
        else if (ty == 'Z' || ty == 'H') {
            char b[2] = {(char)ty, ':'}; sputs(out, cap, l, b, 2);
            const uint8_t *z = memchr(q, 0, (size_t)(end - q)); if (!z) return -1;
            sputs(out, cap, l, (const char *)q, (size_t)(z - q)); q = z + 1;
        } else if (ty == 'B') {
            if (end - q < 5) return -1;
            uint8_t st = q[0]; uint32_t cnt = le32(q + 1); q += 5;
            size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
            if (!es || (uint64_t)(end - q) < (uint64_t)cnt * es) return -1;
            char b[3] = {'B', ':', (char)st}; sputs(out, cap, l, b, 3);
            for (uint32_t i = 0; i < cnt; i++) {
                sputs(out, cap, l, ",", 1);
                if (st == 'f') { float x; memcpy(&x, q, 4); sprintf_g(out, cap, l, x); }
                else if (st == 'c') sprintf_i(out, cap, l, (int8_t)q[0]); else if (st == 'C') sprintf_i(out, cap, l, q[0]);
                else if (st == 's') sprintf_i(out, cap, l, (int16_t)le16(q)); else if (st == 'S') sprintf_i(out, cap, l, le16(q));
                else if (st == 'i') sprintf_i(out, cap, l, (int32_t)le32(q)); else sprintf_i(out, cap, l, le32(q));
                q += es;
            }
        } else return -1;
    }
    return sputs(out, cap, l, "\n", 1);
}

/* ============================================================ writer */
/* one BGZF block (gzip member with the BC extra field) of n <= BGZF_BLOCK payload bytes into out (>= n + 64 + n/1000 bytes);
 * returns its size, 0 on error.  Level 0 ("-u") writes the stored deflate block by hand: no zlib state at all.  zs: a
 * deflate stream to reuse (deflateReset), or NULL to set one up for this block. */
#define BGZF_OUT_STRIDE (65536 + 1024)
static size_t bgzf_pack_block(int level, z_stream *zs, const uint8_t *data, size_t n, uint8_t *out)
{
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    size_t clen;
    memcpy(out, hdr, 16);
    if (level == 0) {
        out[18] = 1;                                            /* BFINAL = 1, BTYPE = 00 (stored) */
        put16(out + 19, (uint32_t)n); put16(out + 21, (uint32_t)(~n & 0xffffu));
        memcpy(out + 23, data, n);
        clen = 5 + n;
    } else {
        z_stream local; int own = 0;
        if (!zs) {
            memset(&local, 0, sizeof local);
            if (deflateInit2(&local, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
            zs = &local; own = 1;
        } else if (deflateReset(zs) != Z_OK) return 0;
        zs->next_in = (Bytef *)data; zs->avail_in = (uInt)n;
        zs->next_out = out + 18; zs->avail_out = BGZF_OUT_STRIDE - 18 - 8;
        int rc = deflate(zs, Z_FINISH);
        clen = BGZF_OUT_STRIDE - 18 - 8 - zs->avail_out;
        if (own) deflateEnd(zs);
        if (rc != Z_STREAM_END) return 0;
    }
    put16(out + 16, (uint32_t)(clen + 18 + 8 - 1));
    put32(out + 18 + clen, crc32x(0, data, n));
    put32(out + 18 + clen + 4, (uint32_t)n);
    return clen + 26;
}

static int bgzf_flush_block(bio_file *f, const uint8_t *data, size_t n)
{
    uint8_t out[BGZF_OUT_STRIDE];
    if (f->w_level != 0 && !f->wz_init) {
        memset(&f->wzs, 0, sizeof f->wzs);
        if (deflateInit2(&f->wzs, f->w_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
        f->wz_init = 1;
    }
    const size_t k = bgzf_pack_block(f->w_level, f->w_level ? &f->wzs : NULL, data, n, out);
    if (!k) return -1;
    return fwrite(out, 1, k, f->fp) == k ? 0 : -1;
}

/* ---- bulk output: whole blocks of a large buffer are packed (deflate / stored + CRC) on worker threads, written in order */
#define BGZF_STORED_SIZE (BGZF_BLOCK + 31)     /* a full block of "-u" output: 18 header + 5 stored-block header + payload + 8 trailer */
typedef struct { int level; const uint8_t *data; size_t nblk; uint8_t *out; size_t *olen; int id, nthr, err; size_t stride; } bgzf_wjob;
static void *bgzf_wworker(void *arg);
static void *bgzf_wworker_spawned(void *arg) { worker_step_back(); return bgzf_wworker(arg); }
static void *bgzf_wworker(void *arg)
{
    bgzf_wjob *j = arg;
    z_stream zs; int have = 0;
    if (j->level != 0) { memset(&zs, 0, sizeof zs); if (deflateInit2(&zs, j->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { j->err = 1; return NULL; } have = 1; }
    for (size_t k = (size_t)j->id; k < j->nblk; k += (size_t)j->nthr) {
        j->olen[k] = bgzf_pack_block(j->level, have ? &zs : NULL, j->data + k * BGZF_BLOCK, BGZF_BLOCK, j->out + k * j->stride);
        if (!j->olen[k]) { j->err = 1; break; }
    }
    if (have) deflateEnd(&zs);
    return NULL;
}

static int w_bytes(bio_file *f, const uint8_t *p, size_t n)
{
    if (!f->w_bam) return fwrite(p, 1, n, f->fp) == n ? 0 : -1;
    while (n) {
        size_t k = BGZF_BLOCK - f->wlen; if (k > n) k = n;
        memcpy(f->wbuf + f->wlen, p, k); f->wlen += k; p += k; n -= k;
        if (f->wlen == BGZF_BLOCK) { if (bgzf_flush_block(f, f->wbuf, f->wlen)) return -1; f->wlen = 0; }
    }
    return 0;
}

int bio_write_raw(bio_file *f, const uint8_t *p, size_t n)
{   /* BAM output: append n bytes of whole records (see bamio.h) */
    if (!f->w_bam) return -1;
    if (f->wlen) {                                           /* complete the pending block first */
        size_t k = BGZF_BLOCK - f->wlen; if (k > n) k = n;
        if (w_bytes(f, p, k)) return -1;
        p += k; n -= k;
    }
    enum { BATCH = 256 };
    int nthr = f->threads > 1 ? f->threads : 1;
    if (f->w_level == 0 && nthr > 4) nthr = 4;               /* stored blocks are a memcpy + CRC32: four threads outrun any pipe or disk */
    /* stored blocks all have the same size: packed back to back, a batch leaves in one write */
    const size_t stride = f->w_level == 0 ? BGZF_STORED_SIZE : BGZF_OUT_STRIDE;
    if (n >= BGZF_BLOCK && !f->wout) {
        f->wout = malloc((size_t)2 * BATCH * BGZF_OUT_STRIDE); f->wolen = malloc(sizeof(size_t) * 2 * BATCH);
        if (!f->wout || !f->wolen) return -1;
    }
    /* two batch buffers: while this thread writes batch k (to a pipe that is the slow part), the workers pack batch k + 1 */
    uint8_t *prev_out = NULL; size_t *prev_len = NULL; size_t prev_blk = 0; int which = 0, rc = 0;
    while (n >= BGZF_BLOCK || prev_out) {
        size_t nblk = n / BGZF_BLOCK; if (nblk > BATCH) nblk = BATCH;
        uint8_t *out = f->wout + (size_t)which * BATCH * BGZF_OUT_STRIDE; size_t *olen = f->wolen + (size_t)which * BATCH;
        int t = nthr; if ((size_t)t > nblk) t = (int)nblk;
        pthread_t th[64]; bgzf_wjob job[64]; int spawned[64];
        const int overlap = prev_out != NULL && nblk > 0;       /* something to write meanwhile: every packing job goes to a spawned thread */
        for (int i = 0; i < t; i++) {
            job[i] = (bgzf_wjob){ f->w_level, p, nblk, out, olen, i, t, 0, stride };
            spawned[i] = (i || overlap) && pthread_create(&th[i], NULL, bgzf_wworker_spawned, &job[i]) == 0;
        }
        if (prev_out && !rc) {
            if (f->w_level == 0) { if (fwrite(prev_out, 1, prev_blk * stride, f->fp) != prev_blk * stride) rc = -1; }
            else for (size_t k = 0; k < prev_blk && !rc; k++)
                if (fwrite(prev_out + k * BGZF_OUT_STRIDE, 1, prev_len[k], f->fp) != prev_len[k]) rc = -1;
        }
        for (int i = 0; i < t; i++) { if (spawned[i]) pthread_join(th[i], NULL); else bgzf_wworker(&job[i]); if (job[i].err) rc = -1; }
        prev_out = nblk ? out : NULL; prev_len = olen; prev_blk = nblk; which ^= 1;
        p += nblk * BGZF_BLOCK; n -= nblk * BGZF_BLOCK;
        if (rc) return -1;
    }
    return n ? w_bytes(f, p, n) : 0;
}

bio_file *bio_open_write(const char *path, const char *mode)
{
    bio_file *f = calloc(1, sizeof *f);
    if (!f) return NULL;
    f->writing = 1;
    f->w_bam = strchr(mode, 'b') != NULL;
    f->w_header = f->w_bam || strchr(mode, 'h') != NULL;
    f->w_level = strchr(mode, 'u') ? 0 : Z_DEFAULT_COMPRESSION;
    if (strcmp(path, "-") == 0) { f->fp = stdout; widen_pipe(stdout); } else { f->fp = fopen(path, "wb"); f->own_fp = 1; }
    if (!f->fp) { free(f); return NULL; }
    if (f->w_bam) f->wbuf = malloc(BGZF_BLOCK);
    return f;
}

int bio_write_header(bio_file *f, const bio_hdr *h)
{
    if (!f->w_header) return 0;
    if (!f->w_bam) return fwrite(h->text, 1, h->l_text, f->fp) == h->l_text && fflush(f->fp) == 0 ? 0 : -1;
    /* the binary header in one buffer, through the bulk writer (blocks packed on the worker threads: the header of a 1 M-sequence
       catalogue is 40 MB) -- the same bytes as writing it field by field */
    size_t total = 12 + h->l_text;
    for (int32_t i = 0; i < h->n_targets; i++) total += 8 + strlen(h->target_name[i]) + 1;
    uint8_t *buf = malloc(total), *p = buf;
    if (!buf) return -1;
    memcpy(p, "BAM\1", 4); put32(p + 4, (uint32_t)h->l_text); p += 8;
    memcpy(p, h->text, h->l_text); p += h->l_text;
    put32(p, (uint32_t)h->n_targets); p += 4;
    for (int32_t i = 0; i < h->n_targets; i++) {
        const size_t ln = strlen(h->target_name[i]) + 1;
        put32(p, (uint32_t)ln); memcpy(p + 4, h->target_name[i], ln); put32(p + 4 + ln, h->target_len[i]);
        p += 8 + ln;
    }
    const int wrc = bio_write_raw(f, buf, total);
    free(buf);
    if (wrc) return -1;
    if (f->wlen) { if (bgzf_flush_block(f, f->wbuf, f->wlen)) return -1; f->wlen = 0; }   /* header in its own block(s), as htslib */
    /* nothing of the header may wait in the stdio buffer for the first records: a downstream process parses the header (and
       starts up its own device) while this one works on its first chunk */
    return fflush(f->fp) ? -1 : 0;
}

int bio_write_record(bio_file *f, const bio_hdr *h, const uint8_t *rec, size_t len)
{
    if (f->w_bam) return w_bytes(f, rec, len);
    size_t l = 0;
    if (bio_sam_format(rec, len, h, &f->fmt, &f->fmt_cap, &l)) return -1;
    return fwrite(f->fmt, 1, l, f->fp) == l ? 0 : -1;
}

int bio_close(bio_file *f)
{
    int rc = 0;
    if (!f) return 0;
    if (f->writing) {
        if (f->w_bam) {
            if (f->wlen && bgzf_flush_block(f, f->wbuf, f->wlen)) rc = -1;
            static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (fwrite(eof, 1, 28, f->fp) != 28) rc = -1;
        }
        if (fflush(f->fp)) rc = -1;
        if (f->wz_init) deflateEnd(&f->wzs);
    } else { if (f->z_init) inflateEnd(&f->zs); io_stop(f); }
    if (f->own_fp && fclose(f->fp)) rc = -1;
    free(f->in); free(f->dec); free(f->line); free(f->ht); free(f->wbuf); free(f->wout); free(f->wolen); free(f->fmt); free(f);
    return rc;
}

char *bio_stringify_argv(int argc, char *argv[])
{
    size_t n = 1;
    for (int i = 0; i < argc; i++) n += strlen(argv[i]) + 1;
    char *s = malloc(n), *p = s;
    if (!s) return NULL;
    for (int i = 0; i < argc; i++) {
        if (i) *p++ = ' ';
        for (const char *c = argv[i]; *c; c++) *p++ = (*c == '\t') ? ' ' : *c;
    }
    *p = 0;
    return s;
}
