/*
 * recwalk.h -- multi-threaded walk of the block_size chain of an uncompressed BAM record stream (the host offset index
 * of include/msamtools_b200.h; sam_read1's framing, msam_helper.c:267).
 *
 * The chain is sequential by nature (every record says where the next one starts) and costs a cache miss per record on
 * freshly inflated data -- about 0.1 us, a second per 10 M records, as much as inflating them on 16 threads.  Here the
 * byte range is cut into one segment per thread; every thread but the first GUESSES a record start at the head of its
 * segment (a run of consecutive plausible record headers) and walks from there.  The guesses are then VERIFIED: the
 * chain of segment t, which is the true chain when its own start is true, must arrive exactly at the start guessed
 * for segment t+1; from the first segment whose guess it misses the rest is walked again sequentially.  So the result
 * is the sequential walk's by construction, whatever the heuristic does.
 */
#ifndef MSG_RECWALK_H
#define MSG_RECWALK_H
#include <stddef.h>
#include <stdint.h>

/* Index the whole records in raw[from, len).  off[0 .. *n] are the offsets known so far (off[*n] == from is the start of
 * the first unindexed record); new record END offsets are appended, *n grows, *off is realloc'ed (capacity *cap entries).
 * A partial trailing record is left unindexed.  n_targets bounds refID in the guess (<= 0: not used).
 * Returns 0, -1 corrupt record (block_size < 32 or > 2^31-1 on the true chain), -2 out of memory. */
int rw_index(const uint8_t *raw, size_t from, size_t len, int32_t n_targets, int threads,
             uint64_t **off, size_t *n, size_t *cap);

#endif
