/* test harness for crc32x.c: compares with zlib's crc32 on every length 0..4100 at four alignments, on a chained CRC over random
 * piece sizes, and on one large buffer; prints "ok <accelerated>" or the first mismatches */
#include <stdio.h>
#include <stdlib.h>
#include <zlib.h>
#include "crc32x.h"

int main(void)
{
    const size_t N = (size_t)9 << 20;
    unsigned char *b = malloc(N + 64);
    unsigned s = 12345;
    for (size_t i = 0; i < N + 64; i++) { s = s * 1103515245u + 12345u; b[i] = (unsigned char)(s >> 16); }
    int bad = 0;
    for (size_t len = 0; len <= 4100; len++)
        for (int al = 0; al < 4; al++) {
            const unsigned seed = (unsigned)(len * 2654435761u);
            const uint32_t a = (uint32_t)crc32(seed, b + al * 5, (uInt)len), c = crc32x(seed, b + al * 5, len);
            if (a != c && bad++ < 5) printf("mismatch: len %zu alignment %d: zlib %08x, crc32x %08x\n", len, al * 5, a, c);
        }
    uint32_t ra = 0, rc = 0;
    for (size_t o = 0; o < N;) {
        s = s * 1103515245u + 12345u;
        size_t k = 1 + (s >> 8) % 70000; if (o + k > N) k = N - o;
        ra = (uint32_t)crc32(ra, b + o, (uInt)k); rc = crc32x(rc, b + o, k); o += k;
    }
    if (ra != rc) { printf("chained mismatch %08x %08x\n", ra, rc); bad++; }
    if ((uint32_t)crc32(0, b + 3, (uInt)N) != crc32x(0, b + 3, N)) { printf("large-buffer mismatch\n"); bad++; }
    if (crc32x(7, NULL, 0) != (uint32_t)crc32(7, Z_NULL, 0)) { printf("NULL convention differs\n"); bad++; }
    if (!bad) printf("ok %d\n", crc32x_accelerated());
    free(b);
    return bad != 0;
}
