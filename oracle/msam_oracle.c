/*
 * msam_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see msam_oracle.h).
 *
 * Sequential restatement of msamtools v1.1.3 (reference at /root/reference):
 *   msam_filter.c:31-35,98-263   filter predicates, pool loop, best-hit writers
 *   mBamVector.c:23-133          bam_cigar2details, bam_get_summary (MD tokenizer)
 *   msam_profile.c:65-425        pool counting, proportional sharing
 *   msam_coverage.c:33-139,189-219
 * plus the htslib 1.24 behaviour those call sites depend on (aux walk,
 * bam_aux2i typing, kstrtok empty tokens), restated from the SAM spec.
 */
#include "msam_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ BAM record view (SAM spec 4.2) */
typedef struct {
    const uint8_t *p;        /* block_size field */
    int32_t  block_size, tid, pos, l_seq;
    uint32_t l_qname, n_cigar, flag;
    const char    *qname;
    const uint8_t *cigar;    /* unaligned little-endian u32[n_cigar] */
    const uint8_t *aux, *end;
} rec_t;

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

static int rec_view(const uint8_t *raw, const uint64_t *off, size_t i, rec_t *r)
{
    const uint8_t *p = raw + off[i];
    uint64_t len = off[i + 1] - off[i];
    r->p = p;
    if (len < 36) return ORC_EFORMAT;
    r->block_size = (int32_t)le32(p);
    r->tid = (int32_t)le32(p + 4);
    r->pos = (int32_t)le32(p + 8);
    r->l_qname = p[12];
    r->n_cigar = le16(p + 16);
    r->flag = le16(p + 18);
    r->l_seq = (int32_t)le32(p + 20);
    r->qname = (const char *)p + 36;
    r->cigar = p + 36 + r->l_qname;
    r->end = p + len;
    if (r->l_seq < 0) return ORC_EFORMAT;
    r->aux = r->cigar + 4 * (uint64_t)r->n_cigar + ((uint64_t)r->l_seq + 1) / 2 + (uint64_t)r->l_seq;
    if (r->aux > r->end) return ORC_EFORMAT;
    return ORC_OK;
}

/* htslib bam_aux_get: first TLV whose tag matches; returns pointer to the type byte.
 * Stops (NULL) at the first malformed / unknown-typed field. */
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end)
{   /* s at type byte; returns pointer past the value or NULL */
    if (s >= end) return NULL;
    uint8_t t = *s++;
    size_t sz;
    switch (t) {
    case 'A': case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    case 'd': sz = 8; break;
    case 'Z': case 'H':
        while (s < end && *s) s++;
        return s < end ? s + 1 : NULL;
    case 'B': {
        if (end - s < 5) return NULL;
        uint8_t st = *s; uint32_t n = le32(s + 1); size_t es;
        switch (st) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break;
                      case 'i': case 'I': case 'f': es = 4; break; default: return NULL; }
        s += 5;
        if ((uint64_t)(end - s) < (uint64_t)n * es) return NULL;
        return s + (size_t)n * es;
    }
    default: return NULL;
    }
    if ((size_t)(end - s) < sz) return NULL;
    return s + sz;
}

static const uint8_t *aux_get(const rec_t *r, const char tag[2])
{
    const uint8_t *s = r->aux;
    while (s && r->end - s >= 3) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) {
            /* htslib validates that the value fits before returning it */
            return aux_skip(s + 2, r->end) ? s + 2 : NULL;
        }
        s = aux_skip(s + 2, r->end);
    }
    return NULL;
}

/* htslib bam_aux2i: integer types only, anything else -> 0 */
static int64_t aux2i(const uint8_t *s)
{
    switch (*s++) {
    case 'c': return (int8_t)s[0];
    case 'C': return s[0];
    case 's': return (int16_t)le16(s);
    case 'S': return le16(s);
    case 'i': return (int32_t)le32(s);
    case 'I': return le32(s);
    default:  return 0;
    }
}

/* ------------------------------------------------------------------ mBamVector.c:23-38 */
static void cigar2details(const rec_t *r, int32_t *alen, int32_t *qlen, int32_t *qclip)
{
    *alen = *qlen = *qclip = 0;
    for (uint32_t k = 0; k < r->n_cigar; k++) {
        uint32_t c = le32(r->cigar + 4 * k);
        int op = c & 0xf; int w = (int)(c >> 4);
        if (op == 5 /*H*/ || op == 4 /*S*/) { *qclip += w; *qlen += w; }
        else if (!(op == 3 /*N*/ || op == 6 /*P*/)) {
            *alen += w;
            if (op == 0 || op == 7 || op == 8 || op == 1) *qlen += w;
        }
    }
}

/* ------------------------------------------------------------------ mBamVector.c:40-133 */
static void get_summary(const rec_t *r, const uint8_t *mdz, int32_t *alen_o, int32_t *qlen_o, int32_t *qclip_o, int32_t *edit_o)
{
    int32_t alen = 0, qlen = 0, qclip = 0, edit = 0;
    for (uint32_t k = 0; k < r->n_cigar; k++) {
        uint32_t c = le32(r->cigar + 4 * k);
        int op = c & 0xf; int w = (int)(c >> 4);
        switch (op) {
        case 0: case 7: case 8: qlen += w; alen += w; break;       /* M = X   :65-71 */
        case 1: qlen += w; /* fall through */                      /* I       :74-76 */
        case 2: edit += w; alen += w; break;                       /* D       :79-82 */
        case 5: case 4: qclip += w; qlen += w; break;              /* H S     :85-89 */
        default: break;                                            /* N P ... :92-95 */
        }
    }
    /* :112-118 -- kstrtok(md, "^0123456789") yields every maximal run of
     * non-delimiter bytes, empty ones included; a run adds its length to edit
     * iff it does not start the string and the byte before it is not '^'. */
    if (mdz && (*mdz == 'Z' || *mdz == 'H')) {
        const char *md = (const char *)mdz + 1;
        const char *p = md;
        for (;;) {
            const char *q = p;
            while (*q && !(*q == '^' || (*q >= '0' && *q <= '9'))) q++;
            if (p > md && p[-1] != '^') edit += (int32_t)(q - p);
            if (!*q) break;
            p = q + 1;
        }
    }
    *alen_o = alen; *qlen_o = qlen; *qclip_o = qclip; *edit_o = edit;
}

/* msam_filter.c:145-157: MD wins over NM; returns 2 (MD), 1 (NM) or 0 (neither) */
static int alignment_stats(const rec_t *r, int32_t *alen, int32_t *qlen, int32_t *qclip, int32_t *edit)
{
    const uint8_t *md = aux_get(r, "MD");
    if (md) { get_summary(r, md, alen, qlen, qclip, edit); return 2; }
    const uint8_t *nm = aux_get(r, "NM");
    if (!nm) { *alen = *qlen = *qclip = *edit = 0; return 0; }
    cigar2details(r, alen, qlen, qclip);
    *edit = (int32_t)aux2i(nm);
    return 1;
}

int orc_record_stats(const uint8_t *raw, const uint64_t *off, size_t n,
                     int32_t *alen, int32_t *qlen, int32_t *qclip, int32_t *edit,
                     int32_t *score, uint8_t *has_as, uint8_t *has_tag)
{
    for (size_t i = 0; i < n; i++) {
        rec_t r; int rc = rec_view(raw, off, i, &r);
        if (rc) return rc;
        int32_t a, q, c, e;
        int t = alignment_stats(&r, &a, &q, &c, &e);
        if (alen) alen[i] = a;
        if (qlen) qlen[i] = q;
        if (qclip) qclip[i] = c;
        if (edit) edit[i] = e;
        if (has_tag) has_tag[i] = (uint8_t)t;
        const uint8_t *as = aux_get(&r, "AS");
        if (has_as) has_as[i] = as != NULL;
        if (score) score[i] = as ? (int32_t)aux2i(as) : 0;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ msam_filter.c:31-35 */
static int filter_fails(const orc_filter_cfg *g, int32_t length, int32_t qlen, int32_t qclip, int32_t edit)
{
    /* dispatch table :73,79-81 : a predicate participates only when its option is active */
    if (g->min_length > 0 && length < g->min_length) return 1;                 /* _FILTER_L */
    if (g->ppt != 0) {                                                         /* _FILTER_P */
        if (g->ppt < 0) { if (1000 * (edit - length) < length * g->ppt) return 1; }
        else            { if (1000 * (length - edit) < length * g->ppt) return 1; }
    }
    if (g->max_clip < 100 && 100 * qclip > g->max_clip * qlen) return 1;       /* _FILTER_Z */
    return 0;
}

typedef struct { uint32_t *e; int32_t *score; uint8_t *has_as; size_t n, cap; } pool_t;

static int pool_push(pool_t *p, uint32_t i, int32_t score, int has_as)
{
    if (p->n == p->cap) {
        size_t nc = p->cap ? 2 * p->cap : 64;       /* pool_limit = 64, doubling: msam_filter.c:103 */
        p->e = realloc(p->e, nc * sizeof *p->e);
        p->score = realloc(p->score, nc * sizeof *p->score);
        p->has_as = realloc(p->has_as, nc);
        if (!p->e || !p->score || !p->has_as) return ORC_ENOMEM;
        p->cap = nc;
    }
    p->e[p->n] = i; p->score[p->n] = score; p->has_as[p->n] = (uint8_t)has_as; p->n++;
    return ORC_OK;
}

typedef struct { const uint8_t *raw; const uint64_t *off; uint32_t *out; size_t n_out; } writer_t;

/* msam_filter.c:206-245 */
static int write_besthit_by_mate(writer_t *w, const pool_t *pool, uint32_t mate_flag, int unique_only)
{
    int best_count = 0; int32_t best_score = INT32_MIN;
    for (size_t i = 0; i < pool->n; i++) {
        uint32_t flag = le16(w->raw + w->off[pool->e[i]] + 18);
        if ((flag & 0xC0) != mate_flag) continue;
        if (!pool->has_as[i]) return ORC_ENOAS;                      /* :219-221 */
        int32_t s = pool->score[i];
        if (s > best_score) { best_score = s; best_count = 1; }
        else if (s == best_score) best_count++;
    }
    if (best_count == 0 || (unique_only && best_count != 1)) return ORC_OK;
    for (size_t i = 0; i < pool->n; i++) {
        uint32_t flag = le16(w->raw + w->off[pool->e[i]] + 18);
        if ((flag & 0xC0) != mate_flag) continue;
        if (pool->score[i] == best_score) w->out[w->n_out++] = pool->e[i];
    }
    return ORC_OK;
}

/* msam_filter.c:196-204,247-263 and mWriteBamPool mBamVector.c:342-347 */
static int write_pool(writer_t *w, const pool_t *pool, int hit_mode)
{
    if (hit_mode == 0) {
        for (size_t i = 0; i < pool->n; i++) w->out[w->n_out++] = pool->e[i];
        return ORC_OK;
    }
    int paired = 0;
    for (size_t i = 0; i < pool->n; i++)
        if (le16(w->raw + w->off[pool->e[i]] + 18) & 0xC0) { paired = 1; break; }
    int uniq = hit_mode == 2, rc;
    if (paired) {
        if ((rc = write_besthit_by_mate(w, pool, 0x40, uniq))) return rc;
        return write_besthit_by_mate(w, pool, 0x80, uniq);
    }
    return write_besthit_by_mate(w, pool, 0, uniq);
}

/* msam_filter.c:98-190 */
int orc_filter(const uint8_t *raw, const uint64_t *off, size_t n, const orc_filter_cfg *g,
               uint32_t *out_idx, size_t *n_out)
{
    writer_t w = { raw, off, out_idx, 0 };
    if (!g->do_filter) {             /* stage disabled: identity stream */
        for (size_t i = 0; i < n; i++) out_idx[i] = (uint32_t)i;
        *n_out = n; return ORC_OK;
    }
    int has_filter = (g->min_length > 0) || (g->ppt != 0) || (g->max_clip < 100);   /* :79-85 */
    int need_stats = has_filter || g->rescore;                                      /* :104 */
    pool_t pool = { 0 };
    char prev_read[256]; prev_read[0] = 0;
    int rc = ORC_OK;
    for (size_t i = 0; i < n; i++) {
        rec_t r;
        if ((rc = rec_view(raw, off, i, &r))) break;
        if (prev_read[0] != 0 && strcmp(r.qname, prev_read) != 0) {                 /* :120-125 */
            if ((rc = write_pool(&w, &pool, g->hit_mode))) break;
            pool.n = 0;
        }
        if (r.flag & 4) {                                                           /* :132-138 */
            if (has_filter && g->keep_unmapped) {
                if (g->ppt >= 0 && g->invert == 1) {
                    const uint8_t *as = aux_get(&r, "AS");
                    if ((rc = pool_push(&pool, (uint32_t)i, as ? (int32_t)aux2i(as) : 0, as != NULL))) break;
                }
            }
            continue;
        }
        int32_t alen = 0, qlen = 0, qclip = 0, edit = 0;
        if (need_stats) {                                                           /* :145-157 */
            if (!alignment_stats(&r, &alen, &qlen, &qclip, &edit)) { rc = ORC_ENOTAG; break; }
        }
        const uint8_t *as = aux_get(&r, "AS");
        int32_t score = as ? (int32_t)aux2i(as) : 0; int has_as = as != NULL;
        if (g->rescore) { score = (alen - edit) * 1 + edit * -1; has_as = 1; }      /* :160-168 */
        strcpy(prev_read, r.qname);                                                 /* :170 */
        if (!has_filter || filter_fails(g, alen, qlen, qclip, edit) == g->invert) { /* :181-183 */
            if ((rc = pool_push(&pool, (uint32_t)i, score, has_as))) break;
        }
    }
    if (!rc) rc = write_pool(&w, &pool, g->hit_mode);                               /* :186 */
    free(pool.e); free(pool.score); free(pool.has_as);
    *n_out = w.n_out;
    return rc;
}

/* sam_write1 of a BAM body; --rescore: bam_aux_del(first AS) + bam_aux_append("AS",'i') :160-168 */
int orc_emit_records(const uint8_t *raw, const uint64_t *off, const uint32_t *idx, size_t m,
                     const orc_filter_cfg *g, uint8_t *out, size_t cap, size_t *nbytes)
{
    size_t o = 0;
    for (size_t j = 0; j < m; j++) {
        rec_t r; int rc = rec_view(raw, off, idx[j], &r);
        if (rc) return rc;
        size_t len = (size_t)(r.end - r.p);
        int do_rescore = g->do_filter && g->rescore && !(r.flag & 4);
        if (!do_rescore) {
            if (o + len > cap) return ORC_EFORMAT;
            memcpy(out + o, r.p, len); o += len; continue;
        }
        int32_t alen, qlen, qclip, edit;
        if (!alignment_stats(&r, &alen, &qlen, &qclip, &edit)) return ORC_ENOTAG;
        int32_t score = (alen - edit) - edit;
        const uint8_t *as = aux_get(&r, "AS");
        size_t cut0 = 0, cut1 = 0;
        if (as) { cut0 = (size_t)(as - 2 - r.p); cut1 = (size_t)(aux_skip(as, r.end) - r.p); }
        size_t nlen = len - (cut1 - cut0) + 7;
        if (o + nlen > cap) return ORC_EFORMAT;
        uint8_t *d = out + o;
        memcpy(d, r.p, as ? cut0 : len);
        size_t k = as ? cut0 : len;
        if (as) { memcpy(d + k, r.p + cut1, len - cut1); k += len - cut1; }
        d[k++] = 'A'; d[k++] = 'S'; d[k++] = 'i';
        d[k++] = (uint8_t)score; d[k++] = (uint8_t)(score >> 8); d[k++] = (uint8_t)(score >> 16); d[k++] = (uint8_t)(score >> 24);
        uint32_t bs = (uint32_t)(nlen - 4);
        d[0] = (uint8_t)bs; d[1] = (uint8_t)(bs >> 8); d[2] = (uint8_t)(bs >> 16); d[3] = (uint8_t)(bs >> 24);
        o += nlen;
    }
    *nbytes = o;
    return ORC_OK;
}

/* ------------------------------------------------------------------ profile */
struct orc_profile {
    int32_t n_targets, n_features; int32_t *fmap; int share_type;
    uint32_t *ui; double *d; uint8_t *hit;
    uint32_t uniq, multi, inserts;
    uint64_t *l_off; int32_t *l_fid; size_t n_lists, cap_lists, n_ent, cap_ent;
};

orc_profile *orc_profile_new(int32_t n_targets, int32_t n_features, const int32_t *fmap, int share_type)
{
    orc_profile *p = calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n_targets = n_targets; p->n_features = n_features; p->share_type = share_type;
    p->fmap = malloc(sizeof(int32_t) * (size_t)(n_targets > 0 ? n_targets : 1));
    for (int32_t i = 0; i < n_targets; i++) p->fmap[i] = fmap ? fmap[i] : i;
    size_t nf = (size_t)(n_features > 0 ? n_features : 1);
    p->ui = calloc(nf, sizeof(uint32_t)); p->d = calloc(nf, sizeof(double)); p->hit = calloc(nf, 1);
    p->cap_lists = 1024; p->l_off = malloc(sizeof(uint64_t) * (p->cap_lists + 1)); p->l_off[0] = 0;
    p->cap_ent = 4096; p->l_fid = malloc(sizeof(int32_t) * p->cap_ent);
    return p;
}
void orc_profile_free(orc_profile *p)
{
    if (!p) return;
    free(p->fmap); free(p->ui); free(p->d); free(p->hit); free(p->l_off); free(p->l_fid); free(p);
}

static void list_push(orc_profile *p, const int32_t *f, size_t n)
{
    if (p->n_lists == p->cap_lists) { p->cap_lists *= 2; p->l_off = realloc(p->l_off, sizeof(uint64_t) * (p->cap_lists + 1)); }
    while (p->n_ent + n > p->cap_ent) { p->cap_ent *= 2; p->l_fid = realloc(p->l_fid, sizeof(int32_t) * p->cap_ent); }
    memcpy(p->l_fid + p->n_ent, f, n * sizeof(int32_t));
    p->n_ent += n; p->n_lists++; p->l_off[p->n_lists] = p->n_ent;
}

/* msam_profile.c:65-200; tids[] = core.tid of the pool's records in pool order */
static void count_pool(orc_profile *p, const int32_t *tids, size_t size, int32_t **scratch, size_t *scap)
{
    const int32_t *fmap = p->fmap;
    if (size == 1) { p->ui[fmap[tids[0]]] += 2; p->uniq++; return; }              /* :75-78 */
    if (size == 2) {                                                               /* :80-127 */
        int f0 = fmap[tids[0]], f1 = fmap[tids[1]];
        if (f0 == f1) { p->ui[f0] += 2; p->uniq++; return; }
        p->multi++;
        switch (p->share_type) {
        case 4: break;
        case 1: p->ui[f0] += 2; p->ui[f1] += 2; break;
        case 2: p->ui[f0]++; p->ui[f1]++; break;
        case 3: { int32_t two[2] = { f0, f1 }; list_push(p, two, 2); break; }
        }
        return;
    }
    if (*scap < size) { *scap = 2 * size; *scratch = realloc(*scratch, *scap * sizeof(int32_t)); }
    int32_t *mappers = *scratch; size_t nm = 0;
    for (size_t i = 0; i < size; i++) {                                            /* :136-142 */
        int f = fmap[tids[i]];
        if (!p->hit[f]) { mappers[nm++] = f; p->hit[f] = 1; }
    }
    for (size_t i = 0; i < nm; i++) p->hit[mappers[i]] = 0;                        /* :145 */
    if (nm == 1) { p->ui[mappers[0]] += 2; p->uniq++; return; }                    /* :152-159 */
    p->multi++;                                                                    /* :162 */
    switch (p->share_type) {
    case 4: break;
    case 1: for (size_t i = 0; i < nm; i++) p->ui[mappers[i]] += 2; break;
    case 2: { double share = 1.0 / (int)nm; for (size_t i = 0; i < nm; i++) p->d[mappers[i]] += share; break; }
    case 3: list_push(p, mappers, nm); break;
    }
}

/* msam_profile.c:204-243 */
int orc_profile_push(orc_profile *p, const uint8_t *raw, const uint64_t *off, const uint32_t *idx, size_t m)
{
    char prev_read[256]; prev_read[0] = 0;
    int32_t *tids = NULL; size_t nt = 0, ct = 0;
    int32_t *scratch = NULL; size_t scap = 0;
    for (size_t j = 0; j < m; j++) {
        rec_t r; int rc = rec_view(raw, off, idx ? idx[j] : j, &r);
        if (rc) { free(tids); free(scratch); return rc; }
        if (r.tid == -1) continue;                                                 /* :223-225 */
        if (r.tid < 0 || r.tid >= p->n_targets) { free(tids); free(scratch); return ORC_EFORMAT; }
        if (prev_read[0] != 0 && strcmp(r.qname, prev_read) != 0) {                /* :226-231 */
            count_pool(p, tids, nt, &scratch, &scap);
            nt = 0; p->inserts++;
        }
        strcpy(prev_read, r.qname);
        if (nt == ct) { ct = ct ? 2 * ct : 64; tids = realloc(tids, ct * sizeof(int32_t)); }
        tids[nt++] = r.tid;
    }
    if (nt > 0) { count_pool(p, tids, nt, &scratch, &scap); p->inserts++; }        /* :235-238 */
    free(tids); free(scratch);
    return ORC_OK;
}

int orc_profile_counts(orc_profile *p, uint32_t *ui, double *d)
{
    if (ui) memcpy(ui, p->ui, sizeof(uint32_t) * (size_t)p->n_features);
    if (d) memcpy(d, p->d, sizeof(double) * (size_t)p->n_features);
    return ORC_OK;
}

/* msam_profile.c:248-425 */
int orc_profile_finish(orc_profile *p, double *abundance_out, orc_profile_out *out)
{
    int n = p->n_features;
    double *abundance = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memset(out, 0, sizeof *out);
    for (int i = 0; i < n; i++) abundance[i] = 1.0 * p->ui[i] / 2;                 /* :284-289 */
    uint32_t purged = 0;
    switch (p->share_type) {
    case 4: case 1: memcpy(abundance_out, abundance, sizeof(double) * (size_t)n); break;
    case 2:
        for (int i = 0; i < n; i++) abundance[i] += p->d[i];                       /* :303-308 */
        memcpy(abundance_out, abundance, sizeof(double) * (size_t)n); break;
    case 3: {
        double *a_k = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        double *a_km1 = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        double *inc = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        memcpy(a_k, abundance, sizeof(double) * (size_t)n);
        int k;
        for (k = 1; k < 20; k++) {                                                 /* :331 */
            double delta = 0;
            for (int j = 0; j < n; j++) inc[j] = 0.0f;
            memcpy(a_km1, a_k, sizeof(double) * (size_t)n);
            for (size_t j = 0; j < p->n_lists; j++) {                              /* :341-365 */
                const int32_t *e = p->l_fid + p->l_off[j]; size_t sz = p->l_off[j + 1] - p->l_off[j];
                double sum = 0;
                for (size_t i = 0; i < sz; i++) sum += a_k[e[i]];
                if (sum > 0) for (size_t i = 0; i < sz; i++) inc[e[i]] += (a_k[e[i]] / sum);
            }
            delta = 0;
            for (int j = 0; j < n; j++) {                                          /* :369-379 */
                a_k[j] = abundance[j] + inc[j];
                if (a_k[j] < 1e-20) a_k[j] = 0;
                double diff = a_k[j] - a_km1[j];
                delta += diff * diff;
            }
            delta /= n;                                                            /* :380 */
            out->delta[k - 1] = delta; out->iterations = k;
            if (delta < 1e-10) { out->converged = 1; break; }                      /* :383 */
        }
        memcpy(abundance_out, a_k, sizeof(double) * (size_t)n);
        for (size_t j = 0; j < p->n_lists; j++) {                                  /* :394-404 */
            const int32_t *e = p->l_fid + p->l_off[j]; size_t sz = p->l_off[j + 1] - p->l_off[j];
            double sum = 0;
            for (size_t i = 0; i < sz; i++) sum += a_k[e[i]];
            if (sum == 0) purged++;
        }
        free(a_k); free(a_km1); free(inc);
        break;
    }
    default: free(abundance); return ORC_EFORMAT;
    }
    out->mapped_inserts = p->inserts; out->uniq = p->uniq; out->multi = p->multi; out->purged = purged;
    out->n_lists = p->n_lists; out->n_entries = p->n_ent;
    free(abundance);
    return ORC_OK;
}

/* ------------------------------------------------------------------ coverage */
struct orc_coverage { int32_t n_targets; uint32_t *tlen; int32_t **cov; uint8_t *covered; };

orc_coverage *orc_coverage_new(int32_t n_targets, const uint32_t *target_len)
{
    orc_coverage *c = calloc(1, sizeof *c);
    size_t n = (size_t)(n_targets > 0 ? n_targets : 1);
    c->n_targets = n_targets; c->tlen = malloc(sizeof(uint32_t) * n);
    memcpy(c->tlen, target_len, sizeof(uint32_t) * (size_t)n_targets);
    c->cov = calloc(n, sizeof(int32_t *)); c->covered = calloc(n, 1);
    return c;
}
void orc_coverage_free(orc_coverage *c)
{
    if (!c) return;
    for (int32_t i = 0; i < c->n_targets; i++) free(c->cov[i]);
    free(c->cov); free(c->covered); free(c->tlen); free(c);
}

/* msam_coverage.c:33-87 (pooling, :89-139, is result-neutral: every record gets +1).
 * The reference does not bounds-check pos against tlen (UB); the oracle and the
 * GPU path both ignore bases outside [0, tlen). */
int orc_coverage_push(orc_coverage *c, const uint8_t *raw, const uint64_t *off, const uint32_t *idx, size_t m)
{
    for (size_t j = 0; j < m; j++) {
        rec_t r; int rc = rec_view(raw, off, idx ? idx[j] : j, &r);
        if (rc) return rc;
        if (r.tid < 0) continue;                                                   /* :42 */
        if (r.tid >= c->n_targets) return ORC_EFORMAT;
        if (!c->covered[r.tid]) {                                                  /* :45-49 */
            c->covered[r.tid] = 1;
            c->cov[r.tid] = calloc(c->tlen[r.tid] ? c->tlen[r.tid] : 1, sizeof(int32_t));
        }
        int32_t *cv = c->cov[r.tid]; int64_t pos = r.pos, tl = c->tlen[r.tid];
        for (uint32_t k = 0; k < r.n_cigar; k++) {
            uint32_t cg = le32(r.cigar + 4 * k); int op = cg & 0xf; int w = (int)(cg >> 4);
            switch (op) {
            case 0: case 7: case 8:
                for (int i = 0; i < w; i++) { int64_t q = pos + i; if (q >= 0 && q < tl) cv[q] += 1; }
                pos += w; break;
            case 2: case 3: pos += w; break;
            default: break;
            }
        }
    }
    return ORC_OK;
}

/* msam_coverage.c:189-219 */
int orc_coverage_finish(orc_coverage *c, uint8_t *covered, int64_t *touched, int64_t *sum)
{
    for (int32_t t = 0; t < c->n_targets; t++) {
        covered[t] = c->covered[t]; touched[t] = 0; sum[t] = 0;
        if (!c->covered[t]) continue;
        for (uint32_t i = 0; i < c->tlen[t]; i++) { int32_t v = c->cov[t][i]; touched[t] += (v != 0); sum[t] += v; }
    }
    return ORC_OK;
}
int orc_coverage_depth(orc_coverage *c, int32_t tid, int32_t *depth)
{
    if (tid < 0 || tid >= c->n_targets) return ORC_EFORMAT;
    if (c->covered[tid]) memcpy(depth, c->cov[tid], sizeof(int32_t) * c->tlen[tid]);
    else memset(depth, 0, sizeof(int32_t) * c->tlen[tid]);
    return ORC_OK;
}

/* ------------------------------------------------------------------ whole pipeline (CPU baseline timing) */
int orc_pipeline(const uint8_t *raw, const uint64_t *off, size_t n, const orc_filter_cfg *cfg,
                 int32_t n_targets, int32_t n_features, const int32_t *fmap, int share_type,
                 double *abundance, orc_profile_out *out, size_t *n_kept)
{
    uint32_t *idx = malloc(sizeof(uint32_t) * (n ? n : 1));
    if (!idx) return ORC_ENOMEM;
    size_t m = 0;
    int rc = orc_filter(raw, off, n, cfg, idx, &m);
    if (!rc) {
        orc_profile *p = orc_profile_new(n_targets, n_features, fmap, share_type);
        rc = orc_profile_push(p, raw, off, idx, m);
        if (!rc) rc = orc_profile_finish(p, abundance, out);
        orc_profile_free(p);
    }
    if (n_kept) *n_kept = m;
    free(idx);
    return rc;
}
