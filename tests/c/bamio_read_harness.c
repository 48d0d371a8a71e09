/* Reads a BAM through msamtools_b200/csrc/host/bamio.c with N inflate threads and writes the raw record stream to stdout
 * (tests/test_bamio_threads.py compares it with the stream the file was written from).  usage: harness file.bam threads */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include "bamio.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    bio_file *f = bio_open_read(argv[1]);
    if (!f) { fprintf(stderr, "open failed\n"); return 1; }
    if (atoi(argv[2]) > 0) bio_set_threads(f, atoi(argv[2]));     /* 0: never called = streaming zlib path */
    bio_hdr *h = bio_read_header(f);
    if (!h) { fprintf(stderr, "header: %s\n", bio_error(f)); return 1; }
    uint8_t *buf = NULL; size_t cap = 0, len = 0; int rc;
    while ((rc = bio_read_record(f, h, &buf, &cap, &len)) == 1)
        if (len > (1u << 22)) { fwrite(buf, 1, len, stdout); len = 0; }
    if (rc < 0) { fprintf(stderr, "read: %s\n", bio_error(f)); return 1; }
    fwrite(buf, 1, len, stdout);
    { uint64_t a = 0, b = 0; bio_inflate_stats(f, &a, &b); fprintf(stderr, "fast %llu zlib %llu\n", (unsigned long long)a, (unsigned long long)b); }
    free(buf); bio_hdr_free(h); bio_close(f);
    return 0;
}
