/*
 * msam_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A sequential, single-threaded restatement of the msamtools v1.1.3 hot path
 * (filter statistics -> pool/best-hit writers -> profile -> coverage), each
 * function citing the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; nothing under msamtools_b200/ links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the
 * expected values written in the reference's own test scripts
 * (tests/test_filter.sh, test_besthit.sh, test_profile.sh, test_coverage.sh,
 * test_integration.sh) and, when oracle/_ref is built, against the reference's
 * own object code (oracle/_ref/msamtools).
 *
 * Input layout is the same raw BAM record stream + offset index as the C ABI
 * (include/msamtools_b200.h).
 */
#ifndef MSAM_ORACLE_H
#define MSAM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define ORC_OK       0
#define ORC_ENOTAG  -5   /* same numbering as MSG_E* so tests can compare codes */
#define ORC_ENOAS   -6
#define ORC_EFORMAT -7
#define ORC_ENOMEM  -4

typedef struct orc_filter_cfg {
    int32_t min_length, ppt, max_clip;
    int32_t do_filter, hit_mode, invert, keep_unmapped, rescore;
} orc_filter_cfg;

typedef struct orc_profile_out {
    uint32_t mapped_inserts, uniq, multi, purged;
    int32_t  iterations, converged;
    double   delta[20];
    uint64_t n_lists, n_entries;
} orc_profile_out;

/* per-record alignment summary; has_tag: 0 none, 1 NM path, 2 MD path */
int orc_record_stats(const uint8_t *raw, const uint64_t *off, size_t n,
                     int32_t *alen, int32_t *qlen, int32_t *qclip, int32_t *edit,
                     int32_t *score, uint8_t *has_as, uint8_t *has_tag);

/* mFilterFile + writers: kept record indices in output order; score_out (may be
 * NULL) receives the AS each record carries on output (rescored if asked).    */
int orc_filter(const uint8_t *raw, const uint64_t *off, size_t n, const orc_filter_cfg *cfg,
               uint32_t *out_idx, size_t *n_out);

/* serialise kept records (sam_write1 of BAM bodies), applying --rescore edits */
int orc_emit_records(const uint8_t *raw, const uint64_t *off, const uint32_t *idx, size_t m,
                     const orc_filter_cfg *cfg, uint8_t *out, size_t cap, size_t *nbytes);

/* mEstimateInsertCountOnFile over the stream idx[0..m) (idx NULL = identity).
 * ui/d are accumulated into (caller zeroes).  lists are appended to an
 * internal CSR owned by the handle.                                           */
typedef struct orc_profile orc_profile;
orc_profile *orc_profile_new(int32_t n_targets, int32_t n_features, const int32_t *fmap, int share_type);
void orc_profile_free(orc_profile *p);
int  orc_profile_push(orc_profile *p, const uint8_t *raw, const uint64_t *off,
                      const uint32_t *idx, size_t m);
int  orc_profile_counts(orc_profile *p, uint32_t *ui, double *d);
/* mInsertCountToAbundanceMatrix */
int  orc_profile_finish(orc_profile *p, double *abundance, orc_profile_out *out);

/* mEstimateCoverageOnFile + summary */
typedef struct orc_coverage orc_coverage;
orc_coverage *orc_coverage_new(int32_t n_targets, const uint32_t *target_len);
void orc_coverage_free(orc_coverage *c);
int  orc_coverage_push(orc_coverage *c, const uint8_t *raw, const uint64_t *off,
                       const uint32_t *idx, size_t m);
int  orc_coverage_finish(orc_coverage *c, uint8_t *covered, int64_t *touched, int64_t *sum);
int  orc_coverage_depth(orc_coverage *c, int32_t tid, int32_t *depth);

/* whole pipelines in one call, for timing the CPU baseline */
int orc_pipeline(const uint8_t *raw, const uint64_t *off, size_t n, const orc_filter_cfg *cfg,
                 int32_t n_targets, int32_t n_features, const int32_t *fmap, int share_type,
                 double *abundance, orc_profile_out *out, size_t *n_kept);
#endif
