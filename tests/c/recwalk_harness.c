/* test harness for recwalk.c: recwalk_harness <threads> <n_targets> <from> < raw stream  ->  "rc n" then the offsets, one per line */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "recwalk.h"

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    size_t cap = 1 << 20, len = 0;
    uint8_t *buf = malloc(cap);
    for (;;) { size_t k = fread(buf + len, 1, cap - len, stdin); len += k; if (k == 0) break; if (len == cap) { cap *= 2; buf = realloc(buf, cap); } }
    size_t from = (size_t)atoll(argv[3]), n = 0, ocap = 4;
    uint64_t *off = malloc(ocap * sizeof *off);
    off[0] = from;
    int rc = rw_index(buf, from, len, atoi(argv[2]), atoi(argv[1]), &off, &n, &ocap);
    printf("%d %zu\n", rc, n);
    if (rc == 0) for (size_t i = 0; i <= n; i++) printf("%llu\n", (unsigned long long)off[i]);
    free(off); free(buf);
    return 0;
}
