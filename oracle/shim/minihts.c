/*
 * minihts.c -- the ~25 htslib entry points msamtools v1.1.3 calls, restated over bamio.c.
 * TEST INFRASTRUCTURE ONLY: exists so that the reference's own, unmodified C sources can be
 * compiled into oracle/_ref/msamtools (reference arithmetic, shim I/O -- not htslib 1.24).
 * Behaviours of htslib that the reference's results depend on are kept: bam1_t memory layout
 * (qname NUL-padded to 4 bytes), bam_aux_get first-match walk, bam_aux2i typing, kstrtok
 * returning empty tokens, sam_read1 rejecting records whose CIGAR and SEQ lengths disagree,
 * bin recomputed on read, sam_hdr_add_pg ID uniquing / PP chaining.
 */
#include "htslib/sam.h"
#include "../../msamtools_b200/csrc/host/bamio.h"
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct samFile { bio_file *f; int writing; uint8_t *buf; size_t cap; };

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

samFile *sam_open(const char *fn, const char *mode)
{
    samFile *fp = calloc(1, sizeof *fp);
    if (!fp) return NULL;
    fp->writing = mode[0] == 'w';
    fp->f = fp->writing ? bio_open_write(fn, mode) : bio_open_read(fn);
    if (!fp->f) { free(fp); return NULL; }
    return fp;
}

int sam_close(samFile *fp)
{
    if (!fp) return 0;
    int rc = bio_close(fp->f);
    free(fp->buf); free(fp);
    return rc;
}

static sam_hdr_t *wrap_hdr(bio_hdr *bh)
{
    if (!bh) return NULL;
    sam_hdr_t *h = calloc(1, sizeof *h);
    h->n_targets = bh->n_targets; h->target_len = bh->target_len; h->target_name = bh->target_name; h->priv = bh;
    return h;
}

sam_hdr_t *sam_hdr_read(samFile *fp) { return wrap_hdr(bio_read_header(fp->f)); }
sam_hdr_t *sam_hdr_dup(const sam_hdr_t *h) { return wrap_hdr(bio_hdr_dup((const bio_hdr *)h->priv)); }
void sam_hdr_destroy(sam_hdr_t *h) { if (!h) return; bio_hdr_free((bio_hdr *)h->priv); free(h); }
int sam_hdr_write(samFile *fp, const sam_hdr_t *h) { return bio_write_header(fp->f, (const bio_hdr *)h->priv); }

int sam_hdr_add_pg(sam_hdr_t *h, const char *name, ...)
{
    const char *pn = name, *vn = "", *cl = "", *ds = "";
    va_list ap; va_start(ap, name);
    for (;;) {
        const char *k = va_arg(ap, const char *);
        if (!k) break;
        const char *v = va_arg(ap, const char *);
        if (!strcmp(k, "PN")) pn = v; else if (!strcmp(k, "VN")) vn = v; else if (!strcmp(k, "CL")) cl = v; else if (!strcmp(k, "DS")) ds = v;
    }
    va_end(ap);
    return bio_hdr_add_pg((bio_hdr *)h->priv, name, pn, vn, cl, ds);
}

int sam_hdr_find_tag_hd(sam_hdr_t *h, const char *key, kstring_t *ks)
{
    char *v = bio_hdr_find_hd_tag((const bio_hdr *)h->priv, key);
    if (!v) return -1;
    free(ks->s); ks->s = v; ks->l = strlen(v); ks->m = ks->l + 1;
    return 0;
}

bam1_t *bam_init1(void) { return calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (!b) return; free(b->data); free(b); }

static int ensure(bam1_t *b, size_t n)
{
    if (n <= b->m_data) return 0;
    size_t m = b->m_data ? b->m_data : 64;
    while (m < n) m *= 2;
    uint8_t *d = realloc(b->data, m);
    if (!d) return -1;
    b->data = d; b->m_data = (uint32_t)m;
    return 0;
}

bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src)
{
    if (ensure(dst, (size_t)src->l_data)) return NULL;
    memcpy(dst->data, src->data, (size_t)src->l_data);
    dst->l_data = src->l_data; dst->core = src->core; dst->id = src->id;
    return dst;
}

bam1_t *bam_dup1(const bam1_t *src)
{
    bam1_t *b = bam_init1();
    if (!b) return NULL;
    if (!bam_copy1(b, src)) { bam_destroy1(b); return NULL; }
    return b;
}

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

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b)
{
    size_t len = 0;
    int rc = bio_read_record(fp->f, (const bio_hdr *)h->priv, &fp->buf, &fp->cap, &len);
    if (rc == 0) return -1;
    if (rc < 0) return -2;
    const uint8_t *r = fp->buf;
    bam1_core_t *c = &b->core;
    c->tid = (int32_t)le32(r + 4); c->pos = (int32_t)le32(r + 8);
    uint32_t lq = r[12]; c->qual = r[13]; c->bin = (uint16_t)le16(r + 14);
    c->n_cigar = le16(r + 16); c->flag = (uint16_t)le16(r + 18); c->l_qseq = (int32_t)le32(r + 20);
    c->mtid = (int32_t)le32(r + 24); c->mpos = (int32_t)le32(r + 28); c->isize = (int32_t)le32(r + 32);
    c->l_extranul = (uint8_t)((4 - (lq & 3)) & 3);
    c->l_qname = (uint16_t)(lq + c->l_extranul);
    size_t rest = len - 36 - lq;
    if (ensure(b, (size_t)c->l_qname + rest)) return -2;
    memcpy(b->data, r + 36, lq);
    memset(b->data + lq, 0, c->l_extranul);
    memcpy(b->data + c->l_qname, r + 36 + lq, rest);
    b->l_data = (int)(c->l_qname + rest);
    if ((size_t)c->l_qname + 4 * (size_t)c->n_cigar + ((size_t)c->l_qseq + 1) / 2 + (size_t)c->l_qseq > (size_t)b->l_data) return -4;
    if (c->n_cigar > 0) {       /* htslib: recompute bin, reject CIGAR/SEQ length mismatch */
        const uint32_t *cig = bam_get_cigar(b);
        int64_t rlen = 0, qlen = 0;
        for (uint32_t k = 0; k < c->n_cigar; k++) {
            int op = cig[k] & 0xf; int64_t w = cig[k] >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += w;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += w;
        }
        if ((c->flag & BAM_FUNMAP) || rlen == 0) rlen = 1;
        c->bin = (uint16_t)reg2bin(c->pos, c->pos + rlen);
        int is_bam = bio_is_bam(fp->f);
        if (c->l_qseq > 0 && (!is_bam || !(c->flag & BAM_FUNMAP)) && qlen != c->l_qseq) {
            fprintf(stderr, "[E::sam_read1] CIGAR and query sequence lengths differ for %s\n", bam_get_qname(b));
            return -4;
        }
    }
    return (int)len;
}

int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b)
{
    const bam1_core_t *c = &b->core;
    size_t lq = (size_t)c->l_qname - c->l_extranul;
    size_t total = 36 + lq + ((size_t)b->l_data - c->l_qname);
    if (total > fp->cap) { uint8_t *nb = realloc(fp->buf, total * 2); if (!nb) return -1; fp->buf = nb; fp->cap = total * 2; }
    uint8_t *r = fp->buf;
    put32(r, (uint32_t)(total - 4)); put32(r + 4, (uint32_t)c->tid); put32(r + 8, (uint32_t)(int32_t)c->pos);
    r[12] = (uint8_t)lq; r[13] = c->qual; put16(r + 14, c->bin); put16(r + 16, c->n_cigar); put16(r + 18, c->flag);
    put32(r + 20, (uint32_t)c->l_qseq); put32(r + 24, (uint32_t)c->mtid); put32(r + 28, (uint32_t)(int32_t)c->mpos); put32(r + 32, (uint32_t)(int32_t)c->isize);
    memcpy(r + 36, b->data, lq);
    memcpy(r + 36 + lq, b->data + c->l_qname, (size_t)b->l_data - c->l_qname);
    return bio_write_record(fp->f, (const bio_hdr *)h->priv, r, total) ? -1 : (int)total;
}

/* ---- aux fields */
static uint8_t *aux_skip(uint8_t *s, uint8_t *end)
{   /* s at the type byte; pointer past the value, or NULL if malformed */
    if (s >= end) return NULL;
    uint8_t t = *s++; size_t sz;
    switch (t) {
    case 'A': case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    case 'd': sz = 8; break;
    case 'Z': case 'H': while (s < end && *s) s++; return s < end ? s + 1 : NULL;
    case 'B': {
        if (end - s < 5) return NULL;
        uint8_t st = *s; uint32_t n = le32(s + 1); size_t es;
        switch (st) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break; case 'i': case 'I': case 'f': es = 4; break; default: return NULL; }
        s += 5;
        if ((uint64_t)(end - s) < (uint64_t)n * es) return NULL;
        return s + (size_t)n * es;
    }
    default: return NULL;
    }
    if ((size_t)(end - s) < sz) return NULL;
    return s + sz;
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
    uint8_t *s = bam_get_aux(b), *end = b->data + b->l_data;
    while (s && end - s >= 3) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) {
            if (aux_skip(s + 2, end)) return s + 2;
            errno = EINVAL; return NULL;
        }
        s = aux_skip(s + 2, end);
    }
    errno = ENOENT;
    return NULL;
}

int64_t bam_aux2i(const uint8_t *s)
{
    int type = *s++;
    switch (type) {
    case 'c': return (int8_t)s[0];
    case 'C': return s[0];
    case 's': return (int16_t)le16(s);
    case 'S': return le16(s);
    case 'i': return (int32_t)le32(s);
    case 'I': return le32(s);
    default: errno = EINVAL; return 0;
    }
}

char *bam_aux2Z(const uint8_t *s)
{
    int type = *s++;
    if (type == 'Z' || type == 'H') return (char *)s;
    errno = EINVAL;
    return NULL;
}

int bam_aux_del(bam1_t *b, uint8_t *s)
{
    uint8_t *end = b->data + b->l_data, *p = s - 2;
    uint8_t *next = aux_skip(s, end);
    if (!next) return -1;
    memmove(p, next, (size_t)(end - next));
    b->l_data -= (int)(next - p);
    return 0;
}

int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data)
{
    if (ensure(b, (size_t)b->l_data + 3 + (size_t)len)) return -1;
    uint8_t *p = b->data + b->l_data;
    p[0] = (uint8_t)tag[0]; p[1] = (uint8_t)tag[1]; p[2] = (uint8_t)type;
    memcpy(p + 3, data, (size_t)len);
    b->l_data += 3 + len;
    return 0;
}

hts_pos_t bam_endpos(const bam1_t *b)
{
    hts_pos_t rlen = 0;
    if (!(b->core.flag & BAM_FUNMAP) && b->core.n_cigar > 0) {
        const uint32_t *cig = bam_get_cigar(b);
        for (uint32_t k = 0; k < b->core.n_cigar; k++) { int op = cig[k] & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cig[k] >> 4; }
    }
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

/* htslib kstring.c kstrtok: tokens are the maximal runs between separators, empty ones included */
char *kstrtok(const char *str, const char *sep_in, ks_tokaux_t *aux)
{
    const unsigned char *p, *start, *sep = (const unsigned char *)sep_in;
    if (sep) {
        if (str == 0 && aux->finished) return 0;
        aux->finished = 0;
        if (sep[0] && sep[1]) {
            aux->sep = -1;
            aux->tab[0] = aux->tab[1] = aux->tab[2] = aux->tab[3] = 0;
            for (p = sep; *p; ++p) aux->tab[*p >> 6] |= 1ull << (*p & 0x3f);
        } else aux->sep = sep[0];
    }
    if (aux->finished) return 0;
    else if (str) { start = (const unsigned char *)str; aux->finished = 0; }
    else start = (const unsigned char *)aux->p + 1;
    if (aux->sep < 0) { for (p = start; *p; ++p) if (aux->tab[*p >> 6] >> (*p & 0x3f) & 1) break; }
    else { for (p = start; *p; ++p) if (*p == aux->sep) break; }
    aux->p = (const char *)p;
    if (*p == 0) aux->finished = 1;
    return (char *)start;
}

char *stringify_argv(int argc, char *argv[]) { return bio_stringify_argv(argc, argv); }
