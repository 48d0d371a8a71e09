/*
 * finflate.h -- fast one-shot raw-DEFLATE decoder for BGZF block payloads (see finflate.c).
 */
#ifndef MSG_FINFLATE_H
#define MSG_FINFLATE_H
#include <stddef.h>
#include <stdint.h>

/* Inflate the raw DEFLATE stream in[0, in_len) into exactly out_len bytes at out.
 * Returns 0 on success.  Returns -1 -- leaving out[0, out_len) in an unspecified state, never touching anything outside
 * it -- if the stream is malformed, does not produce exactly out_len bytes, or uses a construct this decoder leaves to
 * zlib; callers fall back to zlib and verify the block CRC in either case.                                              */
int fi_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

#endif
