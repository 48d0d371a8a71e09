/*
 * bamio.h -- from-scratch SAM / BAM / BGZF reader and writer over zlib (C99).
 *
 * htslib is not available in this build (the reference downloads it at build time,
 * deps/htslib/defs.txt:2), so the host side carries its own I/O.  Records always travel
 * as raw BAM records (int32 block_size + body, SAM spec 4.2): that is what the GPU path
 * takes (include/msamtools_b200.h) and what BAM output needs; SAM text is parsed into /
 * printed from that form following htslib's conventions (smallest integer aux type,
 * reg2bin, "*" sequences).
 */
#ifndef MSG_BAMIO_H
#define MSG_BAMIO_H
#include <stddef.h>
#include <stdint.h>

typedef struct bio_hdr {
    char     *text;          /* SAM header text, NUL terminated */
    size_t    l_text;
    int32_t   n_targets;
    char    **target_name;
    uint32_t *target_len;
    char     *name_arena;    /* when set, every target_name[i] points into this one block (a 1 M-sequence catalogue is read,
                                copied and freed as one allocation instead of a million); NULL: names are separate mallocs */
} bio_hdr;

typedef struct bio_file bio_file;

/* ---- reading: "-" is stdin; BGZF/gzip and BAM/SAM are auto-detected (as htslib does) */
bio_file *bio_open_read(const char *path);
bio_hdr  *bio_read_header(bio_file *f);                       /* NULL on error */
/* next record as a raw BAM record appended at (*buf)+(*len); grows *buf. 1 = record, 0 = EOF, <0 = error */
int       bio_read_record(bio_file *f, const bio_hdr *h, uint8_t **buf, size_t *cap, size_t *len);
/* BAM input, bulk: append bytes of the decompressed record stream at buf + *len, never beyond cap.  With worker threads
 * (bio_set_threads) whole BGZF blocks are inflated straight into buf -- no intermediate copy; a record may end up
 * split across two calls, the caller walks the block_size chain itself.  1 = appended, 0 = EOF, -1 = error,
 * 2 = the next block does not fit in cap - *len.  May be mixed with bio_read_record only in the order records-then-raw. */
int       bio_read_raw(bio_file *f, uint8_t *buf, size_t cap, size_t *len);
const char *bio_error(const bio_file *f);
/* BGZF input: read block-wise -- a read-ahead thread, blocks (independent gzip members) inflated by the one-shot decoder of
 * finflate.c on `n` threads (n = 1: the caller's) with zlib as fallback and the CRC checked either way; call before
 * bio_read_header.  Never called: zlib's streaming inflate on the caller's thread (also what plain gzip input always gets).
 * BAM output: threads that pack blocks in bio_write_raw.                                                                    */
void      bio_set_threads(bio_file *f, int n);
/* bytes of decompressed input produced so far / seconds spent producing them (read + inflate)  */
void      bio_ingest_stats(const bio_file *f, uint64_t *bytes, double *seconds);
/* parallel path only: blocks inflated by the fast one-shot decoder (finflate.c) / blocks that went through zlib         */
void      bio_inflate_stats(const bio_file *f, uint64_t *fast_blocks, uint64_t *zlib_blocks);

/* ---- writing: mode "w" SAM, "wh" SAM+header, "wb" BAM, "wbu" BAM in level-0 BGZF (msam_filter.c:464-470) */
bio_file *bio_open_write(const char *path, const char *mode);
int       bio_write_header(bio_file *f, const bio_hdr *h);
int       bio_write_record(bio_file *f, const bio_hdr *h, const uint8_t *rec, size_t len);
/* BAM output, bulk: append n bytes holding whole raw records.  Full BGZF blocks are packed (deflate, or the stored block of
 * "-u", plus CRC32) on the worker threads of bio_set_threads and written in order -- same bytes as record-wise writing.  */
int       bio_write_raw(bio_file *f, const uint8_t *p, size_t n);
int       bio_close(bio_file *f);
int       bio_is_bam(const bio_file *f);

/* ---- header helpers */
bio_hdr *bio_hdr_parse_text(const char *text, size_t l_text); /* builds targets from @SQ lines */
bio_hdr *bio_hdr_dup(const bio_hdr *h);
void     bio_hdr_free(bio_hdr *h);
/* value of TAG on the @HD line (malloc'ed) or NULL */
char    *bio_hdr_find_hd_tag(const bio_hdr *h, const char *tag);
/* append "@PG ID:<unique id> PN:.. [PP:<last in chain>] VN:.. CL:.. DS:.." (htslib sam_hdr_add_pg semantics) */
int      bio_hdr_add_pg(bio_hdr *h, const char *name, const char *pn, const char *vn, const char *cl, const char *ds);
int      bio_hdr_tid(const bio_hdr *h, const char *name);

/* ---- record <-> SAM text */
/* parse one SAM line (no trailing newline needed) into a raw BAM record appended at (*buf)+(*len) */
int      bio_sam_parse(const char *line, size_t l, const bio_hdr *h, uint8_t **buf, size_t *cap, size_t *len, char *err, size_t lerr);
/* format a raw BAM record as a SAM line (with trailing '\n') appended to a growable string */
int      bio_sam_format(const uint8_t *rec, size_t len, const bio_hdr *h, char **out, size_t *cap, size_t *olen);

/* join argv with spaces, quoting nothing -- htslib's stringify_argv (tabs become spaces) */
char    *bio_stringify_argv(int argc, char *argv[]);

#endif
