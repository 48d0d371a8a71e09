/*
 * nulldev.c -- a NULL DEVICE behind include/msamtools_b200.h, for profiling the HOST side of the drop-in CLI
 * (reader thread, BGZF inflate, record index, QNAME split, record output, header and table writers) on a machine
 * without a GPU.  TEST INFRASTRUCTURE ONLY: it computes nothing -- `filter` "keeps" a fixed 4 of 5 records (one memcpy,
 * standing in for gather + D2H), `profile` returns a constant vector -- so its outputs are meaningless and nothing in
 * the product, the tests' parity checks or bench.py's measured legs may load it.  Built by tests/hostprof/run.py into
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
int msg_kept_count(msg_ctx *c, size_t *n) { *n = c->nrec - (c->nrec + 4) / 5; return MSG_OK; }
int msg_pull_records(msg_ctx *c, uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec)
{
    size_t nb = 0, nr = 0;
    for (size_t i = 0; i < c->nrec; i++) {
        if (i % 5 == 0) continue;
        const size_t l = (size_t)(c->off[i + 1] - c->off[i]);
        if (out) { if (nb + l > cap) return MSG_ERANGE; memcpy(out + nb, c->raw + c->off[i], l); }
        nb += l; nr++;
    }
    *nbytes = nb; *nrec = nr;
    return MSG_OK;
}
int msg_finish_profile(msg_ctx *c, double *ab, msg_profile_stats *st)
{
    for (int i = 0; i < c->cfg.n_features; i++) ab[i] = 1.0 + (i % 7) * 0.125;
    memset(st, 0, sizeof *st);
    st->mapped_inserts = (uint32_t)(c->pushed / 3); st->uniq_mapper_count = st->mapped_inserts; st->em_iterations = 1; st->em_converged = 1;
    return MSG_OK;
}
int msg_finish_coverage(msg_ctx *c, uint8_t *cov, int64_t *t, int64_t *s)
{
    for (int i = 0; i < c->cfg.n_targets; i++) { cov[i] = 1; t[i] = 10; s[i] = 20; }
    return MSG_OK;
}
int msg_pull_coverage(msg_ctx *c, int32_t tid, int32_t *depth) { memset(depth, 0, sizeof(int32_t) * c->cfg.target_len[tid]); return MSG_OK; }
int msg_get_timing(msg_ctx *c, msg_timing *t, int reset) { (void)c; (void)reset; memset(t, 0, sizeof *t); return MSG_OK; }

#include "../../msamtools_b200/csrc/host/recindex.c"
