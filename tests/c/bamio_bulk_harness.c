/* Exercises the bulk paths of msamtools_b200/csrc/host/bamio.c (tests/test_bamio_threads.py):
 *   harness read  in.bam threads cap      -- record stream to stdout through bio_read_raw into a fixed buffer of `cap` bytes
 *   harness copy  in.bam threads mode out -- rewrites the file: header, then all records through bio_write_raw (mode "wb"/"wbu")
 *   harness copyr in.bam threads mode out -- the same record by record (bio_write_record), for a byte comparison            */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bamio.h"

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    bio_file *f = bio_open_read(argv[2]);
    if (!f) { fprintf(stderr, "open failed\n"); return 1; }
    if (atoi(argv[3]) > 0) bio_set_threads(f, atoi(argv[3]));     /* 0: never called = streaming zlib path */
    bio_hdr *h = bio_read_header(f);
    if (!h) { fprintf(stderr, "header: %s\n", bio_error(f)); return 1; }
    if (!strcmp(argv[1], "read")) {
        size_t cap = (size_t)atol(argv[4]), len = 0; int rc;
        uint8_t *buf = malloc(cap);
        for (;;) {
            rc = bio_read_raw(f, buf, cap, &len);
            if (rc < 0) { fprintf(stderr, "read: %s\n", bio_error(f)); return 1; }
            if (rc == 0) break;
            if (rc == 2 || len > cap / 2) { fwrite(buf, 1, len, stdout); len = 0; }
        }
        fwrite(buf, 1, len, stdout);
        free(buf);
    } else {
        if (argc < 6) return 2;
        uint8_t *buf = NULL; size_t cap = 0, len = 0; int rc;
        size_t *ends = NULL, n = 0, ncap = 0;
        while ((rc = bio_read_record(f, h, &buf, &cap, &len)) == 1) {
            if (n == ncap) { ncap = ncap ? 2 * ncap : 1024; ends = realloc(ends, ncap * sizeof *ends); }
            ends[n++] = len;
        }
        if (rc < 0) { fprintf(stderr, "read: %s\n", bio_error(f)); return 1; }
        bio_file *o = bio_open_write(argv[5], argv[4]);
        if (!o) return 1;
        bio_set_threads(o, atoi(argv[3]));
        if (bio_write_header(o, h)) return 1;
        if (!strcmp(argv[1], "copy")) {
            /* three uneven pieces, cut on record boundaries, so that pending partial blocks are exercised */
            size_t a = n ? ends[n / 3] : 0, b = n ? ends[(2 * n) / 3] : 0;
            if (bio_write_raw(o, buf, a) || bio_write_raw(o, buf + a, b - a) || bio_write_raw(o, buf + b, len - b)) return 1;
        } else {
            size_t prev = 0;
            for (size_t i = 0; i < n; i++) { if (bio_write_record(o, h, buf + prev, ends[i] - prev)) return 1; prev = ends[i]; }
        }
        if (bio_close(o)) return 1;
        free(buf); free(ends);
    }
    bio_hdr_free(h); bio_close(f);
    return 0;
}
