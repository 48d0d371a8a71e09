/*
 * crc32x.h -- CRC-32 (gzip / BGZF, reflected polynomial 0xEDB88320) with carry-less multiplication where the CPU has it.
 *
 * Every byte that passes through the host side is CRC-checked at least once (BGZF block trailers, SAM spec 4.1; gzip
 * trailers of the tables), kept records of `filter | profile` three times.  zlib's table-driven crc32() runs at ~3 GB/s per
 * thread; folding 64 bytes per iteration with PCLMULQDQ (V. Gopal et al., "Fast CRC Computation for Generic Polynomials
 * Using PCLMULQDQ Instruction", Intel 2009) runs at memory speed.  crc32x() has zlib's calling convention and returns the
 * same values; on CPUs without PCLMULQDQ / SSE4.1 (or other architectures) it IS zlib's crc32().
 */
#ifndef MSG_CRC32X_H
#define MSG_CRC32X_H
#include <stddef.h>
#include <stdint.h>

uint32_t crc32x(uint32_t crc, const void *buf, size_t len);
int crc32x_accelerated(void);      /* 1 if the carry-less path is in use on this machine */

#endif
