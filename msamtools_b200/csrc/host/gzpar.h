/*
 * gzpar.h -- gzip writer whose deflate work runs on worker threads (C99, zlib).
 *
 * The reference writes its tables through gzopen/gzprintf (msam_profile.c:880-1012, mMatrix.c:359-376,
 * msam_coverage.c:143-219): one thread formats and deflates row by row.  For a 1 M-gene catalogue that is seconds per
 * profile; here the text is collected in large blocks, the blocks are deflated concurrently (each primed with the last
 * 32 KB of its predecessor as dictionary and closed with a sync flush, the way pigz does) and written in order as ONE
 * gzip member at zlib's default level -- `zcat`, Python's gzip and pandas read it like gzopen's output, and the
 * decompressed bytes are identical to what the reference's writers produce.
 */
#ifndef MSG_GZPAR_H
#define MSG_GZPAR_H
#include <stddef.h>

typedef struct gzp gzp;

gzp *gzp_open(const char *path, int threads);      /* "-" = stdout; NULL on error */
int  gzp_write(gzp *g, const void *p, size_t n);   /* 0 ok, -1 error */
int  gzp_puts(gzp *g, const char *s);
int  gzp_printf(gzp *g, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
/* room for at least n bytes at the end of the pending text; the caller writes there and calls gzp_commit(written) */
char *gzp_reserve(gzp *g, size_t n);
int  gzp_commit(gzp *g, size_t n);
int  gzp_close(gzp *g);                            /* flushes, writes the trailer, closes; 0 ok */

#endif
