/* test harness for gzpar.c: gzpar_harness <out.gz> <threads> <mode> < input
 *   write   : the input in one gzp_write
 *   pieces  : the input in pieces of 1 .. 70001 bytes through gzp_write / gzp_puts / gzp_printf / gzp_reserve+commit in turn
 *   empty   : nothing at all                                                                                        */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gzpar.h"

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    size_t cap = 1 << 20, n = 0;
    char *buf = malloc(cap);
    for (;;) { size_t k = fread(buf + n, 1, cap - n, stdin); n += k; if (k == 0) break; if (n == cap) { cap *= 2; buf = realloc(buf, cap); } }
    gzp *g = gzp_open(argv[1], atoi(argv[2]));
    if (!g) return 3;
    if (!strcmp(argv[3], "write")) { if (gzp_write(g, buf, n)) return 4; }
    else if (!strcmp(argv[3], "pieces")) {
        size_t o = 0, step = 1; int how = 0;
        while (o < n) {
            size_t k = step > n - o ? n - o : step;
            if (how % 3 == 0 || memchr(buf + o, 0, k)) { if (gzp_write(g, buf + o, k)) return 4; }
            else if (how % 3 == 1) { if (gzp_printf(g, "%.*s", (int)k, buf + o)) return 4; }
            else { char *d = gzp_reserve(g, k + 7); if (!d) return 4; memcpy(d, buf + o, k); if (gzp_commit(g, k)) return 4; }
            o += k; how++;
            step = step * 3 + 1; if (step > 70001) step = 1 + how % 5;
        }
    }
    if (gzp_close(g)) return 5;
    free(buf);
    return 0;
}
