/* Test harness for msamtools_b200/csrc/host/finflate.c (built with -fsanitize=address,undefined by tests/test_finflate.py).
 * Input file: repeated cases  u32 in_len | u32 out_len | u8 must_succeed | in bytes | out_len expected bytes.
 * Every buffer is an exact-size heap block so that the sanitizer sees any access outside it.
 * Output: one line "cases N ok A rejected B" ; exit status 1 on the first wrong answer. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "finflate.h"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    unsigned long n = 0, ok = 0, rej = 0;
    for (;;) {
        uint32_t hdr[2]; uint8_t must;
        if (fread(hdr, 4, 2, fp) != 2) break;
        if (fread(&must, 1, 1, fp) != 1) return 2;
        uint8_t *in = malloc(hdr[0] ? hdr[0] : 1), *exp = malloc(hdr[1] ? hdr[1] : 1), *out = malloc(hdr[1] ? hdr[1] : 1);
        if (hdr[0] && fread(in, 1, hdr[0], fp) != hdr[0]) return 2;
        if (hdr[1] && fread(exp, 1, hdr[1], fp) != hdr[1]) return 2;
        /* shrink to the exact sizes (malloc(1) above only avoids malloc(0)) */
        uint8_t *in2 = malloc(hdr[0] ? hdr[0] : 1); memcpy(in2, in, hdr[0]);
        int rc = fi_inflate(hdr[0] ? in2 : in2, hdr[0], out, hdr[1]);
        n++;
        if (rc == 0) {
            ok++;
            if (memcmp(out, exp, hdr[1]) != 0) { fprintf(stderr, "case %lu: wrong output\n", n); return 1; }
        } else {
            rej++;
            if (must) { fprintf(stderr, "case %lu: rejected a stream it must decode (in %u out %u)\n", n, hdr[0], hdr[1]); return 1; }
        }
        free(in); free(in2); free(exp); free(out);
    }
    printf("cases %lu ok %lu rejected %lu\n", n, ok, rej);
    return 0;
}
