/*
 * nulldev.c -- a NULL DEVICE behind include/msamtools_b200.h, for profiling the HOST side of the drop-in CLI
 * (reader thread, BGZF inflate, record index, QNAME split, record output, header and table writers) on a machine
 * without a GPU, and for testing that plumbing (tests/test_cli_host_nulldev.py, also under ThreadSanitizer).  TEST
 * INFRASTRUCTURE ONLY: it computes nothing -- `filter` "keeps" the records whose 0-based POS is not a multiple of 5 (or is absent) (one memcpy,
 * standing in for gather + D2H), `profile` returns a fixed pattern and the record count, `coverage` a fixed pattern -- so
 * its outputs say nothing about alignments and nothing in the product, the parity tests or bench.py's measured legs may
 * load it.  Built by tests/hostprof/run.py into
 * tests/hostprof/_build/libmsamtools_b200.so and put in front of the real library with LD_LIBRARY_PATH.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/msamtools_b200.h"

struct msg_ctx {
    msg_config cfg;
    const uint8_t *raw; const uint64_t *off; size_t nrec, nbytes;
    uint64_t pushed;
};

int msg_abi_version(void) { return MSG_ABI_VERSION; }
int msg_device_count(void) { return 1; }
const char *msg_last_error(const msg_ctx *c) { (void)c; return "null device"; }
int msg_create(const msg_config *cfg, msg_ctx **out) { msg_ctx *c = calloc(1, sizeof *c); c->cfg = *cfg; *out = c; return MSG_OK; }
void msg_destroy(msg_ctx *c) { free(c); }
int msg_host_alloc(int dev, size_t n, void **p) { (void)dev; return posix_memalign(p, 4096, n) ? MSG_ENOMEM : MSG_OK; }
int msg_host_free(void *p) { free(p); return MSG_OK; }
int msg_push(msg_ctx *c, const uint8_t *raw, size_t nbytes, const uint64_t *off, size_t nrec)
{
    c->raw = raw; c->off = off; c->nrec = nrec; c->nbytes = nbytes; c->pushed += nrec;
    return MSG_OK;
}
int msg_push_async(msg_ctx *c, const uint8_t *raw, size_t nbytes, const uint64_t *off, size_t nrec) { return msg_push(c, raw, nbytes, off, nrec); }
int msg_wait(msg_ctx *c) { (void)c; return MSG_OK; }
int msg_sync(msg_ctx *c) { (void)c; return MSG_OK; }
int msg_reset(msg_ctx *c) { c->pushed = 0; return MSG_OK; }
static int kept(const msg_ctx *c, size_t i)
{
    const uint8_t *r = c->raw + c->off[i];
    const uint32_t pos = (uint32_t)r[8] | (uint32_t)r[9] << 8 | (uint32_t)r[10] << 16 | (uint32_t)r[11] << 24;
    return pos == 0xffffffffu || pos % 5u != 0;           /* (POS 0 = no position is kept, so that unplaced records are exercised) */
}
int msg_kept_count(msg_ctx *c, size_t *n) { size_t k = 0; for (size_t i = 0; i < c->nrec; i++) k += (size_t)kept(c, i); *n = k; return MSG_OK; }
int msg_pull_records(msg_ctx *c, uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec)
{
    size_t nb = 0, nr = 0;
    for (size_t i = 0; i < c->nrec; i++) {
        if (!kept(c, i)) continue;
        const size_t l = (size_t)(c->off[i + 1] - c->off[i]);
        if (out) { if (nb + l > cap) return MSG_ERANGE; memcpy(out + nb, c->raw + c->off[i], l); }
        nb += l; nr++;
    }
    *nbytes = nb; *nrec = nr;
    return MSG_OK;
}
int msg_finish_profile(msg_ctx *c, double *ab, msg_profile_stats *st)
{
    for (int i = 0; i < c->cfg.n_features; i++)
        ab[i] = i % 11 == 0 ? 0.0 : (double)(((uint32_t)i * 2654435761u) >> 7) / 1024.0 / (double)(i % 13 + 1);
    memset(st, 0, sizeof *st);
    st->mapped_inserts = (uint32_t)c->pushed; st->uniq_mapper_count = st->mapped_inserts; st->em_iterations = 1; st->em_converged = 1;
    return MSG_OK;
}
int msg_finish_coverage(msg_ctx *c, uint8_t *cov, int64_t *t, int64_t *s)
{
    for (int i = 0; i < c->cfg.n_targets; i++) { cov[i] = i % 3 != 1; t[i] = 10 + i; s[i] = 20 + 3 * i; }
    return MSG_OK;
}
int msg_pull_coverage(msg_ctx *c, int32_t tid, int32_t *depth)
{
    for (uint32_t i = 0; i < c->cfg.target_len[tid]; i++) depth[i] = (int32_t)((i * 2654435761u + (uint32_t)tid) >> 12) % 100003 - (i % 97 == 0 ? 7 : 0);
    return MSG_OK;
}
int msg_get_timing(msg_ctx *c, msg_timing *t, int reset) { (void)c; (void)reset; memset(t, 0, sizeof *t); return MSG_OK; }

#include "../../msamtools_b200/csrc/host/recindex.c"
