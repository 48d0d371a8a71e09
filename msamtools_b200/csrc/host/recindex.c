/*
 * recindex.c -- host-side record index helpers of the C ABI (include/msamtools_b200.h): msg_index_records walks the
 * block_size chain of an uncompressed BAM record stream (sam_read1's framing, msam_helper.c:267), msg_split_point finds
 * a chunk / shard boundary that no QNAME group straddles (msam_filter.c:120-121,170; msam_profile.c:223-232).
 * Plain C, no CUDA: compiled into libmsamtools_b200.so (the ABI) and into libmsamhost.so, which CPU-only tools -- the
 * bench's reference arm, the sharding helpers -- load without ever mapping the CUDA library.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef MSG_OK
#define MSG_OK       0
#define MSG_EFORMAT -7
#define MSG_ERANGE  -8
#endif

#ifdef __cplusplus
extern "C" {
#endif

int msg_index_records(const uint8_t *raw, size_t nbytes, uint64_t *rec_off, size_t cap, size_t *nrec, size_t *consumed, int allow_partial)
{
    size_t o = 0, n = 0;
    while (o + 4 <= nbytes) {
        uint32_t bs = (uint32_t)raw[o] | (uint32_t)raw[o + 1] << 8 | (uint32_t)raw[o + 2] << 16 | (uint32_t)raw[o + 3] << 24;
        if (bs < 32 || bs > 0x7fffffffu) return MSG_EFORMAT;
        if (o + 4 + (size_t)bs > nbytes) break;
        if (rec_off) { if (n + 1 >= cap) return MSG_ERANGE; rec_off[n] = o; }
        n++; o += 4 + (size_t)bs;
    }
    if (rec_off) { if (n >= cap) return MSG_ERANGE; rec_off[n] = o; }
    if (nrec) *nrec = n;
    if (consumed) *consumed = o;
    if (o != nbytes && !allow_partial) return MSG_EFORMAT;
    return MSG_OK;
}

size_t msg_split_point(const uint8_t *raw, const uint64_t *rec_off, size_t nrec, size_t want)
{
    if (want > nrec) want = nrec;
    if (want == nrec) return nrec;
    for (size_t k = want; k > 0; k--) {
        const uint8_t *a = raw + rec_off[k - 1], *b = raw + rec_off[k];
        uint32_t flag = (uint32_t)a[18] | (uint32_t)a[19] << 8;
        int32_t tid = (int32_t)((uint32_t)a[4] | (uint32_t)a[5] << 8 | (uint32_t)a[6] << 16 | (uint32_t)a[7] << 24);
        if ((flag & 4u) || tid < 0) continue;
        if (a[12] != b[12] || memcmp(a + 36, b + 36, a[12]) != 0) return k;
    }
    return 0;
}

#ifdef __cplusplus
}
#endif
