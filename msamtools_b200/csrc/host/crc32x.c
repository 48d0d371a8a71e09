/* crc32x.c -- see crc32x.h. */
#include "crc32x.h"
#include <zlib.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>

/* Folding constants for the reflected CRC-32 polynomial P(x) = 0x1DB710641 (bit-reflected 0x104C11DB7), as derived in the
 * paper: x^(4*128+32), x^(4*128-32), x^(128+32), x^(128-32), x^64 mod P, then Barrett's mu and P.  tests/test_crc32x.py
 * checks the function against zlib's crc32() on every length 0..4100 at several alignments, on running (chained) CRCs and on
 * large buffers, under ASan/UBSan. */
#define K1 0x0154442bd4ull
#define K2 0x01c6e41596ull
#define K3 0x01751997d0ull
#define K4 0x00ccaa009eull
#define K5 0x0163cd6124ull
#define PX 0x01db710641ull
#define MU 0x01f7011641ull

/* len >= 64 and a multiple of 16; crc is the running (pre-inverted) register */
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(uint32_t crc, const uint8_t *buf, size_t len)
{
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_set_epi64x((long long)K2, (long long)K1);
    buf += 64; len -= 64;
    while (len >= 64) {                                   /* four independent 128-bit lanes, 64 bytes per iteration */
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_set_epi64x((long long)K4, (long long)K3);    /* the four lanes into one */
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i *)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    /* 128 -> 64 bits */
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_set_epi64x(0, (long long)K5);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    /* Barrett reduction to 32 bits */
    x0 = _mm_set_epi64x((long long)MU, (long long)PX);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

static int g_have_clmul;                                  /* decided once, before main() and before any thread exists */
__attribute__((constructor)) static void crc32x_init(void)
{
    __builtin_cpu_init();
    g_have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
}
static int have_clmul(void) { return g_have_clmul; }

int crc32x_accelerated(void) { return have_clmul(); }

uint32_t crc32x(uint32_t crc, const void *buf, size_t len)
{
    const uint8_t *p = buf;
    if (!p) return (uint32_t)crc32(0L, Z_NULL, 0);
    if (len >= 64 && have_clmul()) {
        const size_t n = len & ~(size_t)15;
        crc = ~crc32_clmul(~crc, p, n);
        p += n; len -= n;
    }
    while (len) {                                          /* the remainder (and everything on other machines): zlib */
        const size_t k = len > 0x40000000u ? 0x40000000u : len;
        crc = (uint32_t)crc32(crc, p, (uInt)k);
        p += k; len -= k;
    }
    return crc;
}

#else

int crc32x_accelerated(void) { return 0; }
uint32_t crc32x(uint32_t crc, const void *buf, size_t len)
{
    const uint8_t *p = buf;
    if (!p) return (uint32_t)crc32(0L, Z_NULL, 0);
    while (len) { const size_t k = len > 0x40000000u ? 0x40000000u : len; crc = (uint32_t)crc32(crc, p, (uInt)k); p += k; len -= k; }
    return crc;
}

#endif
