/* Minimal stand-in for htslib/sam.h: exactly the types, macros and functions msamtools v1.1.3
 * uses (SURVEY.md 8c lists them).  TEST INFRASTRUCTURE: lets the reference's own C sources be
 * compiled unchanged into oracle/_ref/msamtools; I/O is msamtools_b200/csrc/host/bamio.c. */
#ifndef SHIM_SAM_H
#define SHIM_SAM_H
#include <stdint.h>
#include <stddef.h>
#include "kstring.h"

typedef int64_t hts_pos_t;

typedef struct sam_hdr_t {
    int32_t n_targets;
    uint32_t *target_len;
    char **target_name;
    void *priv;                      /* bio_hdr* */
} sam_hdr_t;

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual, l_extranul;
    uint16_t flag, l_qname;
    uint32_t n_cigar;
    int32_t l_qseq, mtid;
    hts_pos_t mpos, isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
} bam1_t;

typedef struct samFile samFile;

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_MAX_QNAME_LEN 254         /* the reference expects this from htslib; tests need 254-char names */

#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))

samFile *sam_open(const char *fn, const char *mode);
int sam_close(samFile *fp);
sam_hdr_t *sam_hdr_read(samFile *fp);
int sam_hdr_write(samFile *fp, const sam_hdr_t *h);
sam_hdr_t *sam_hdr_dup(const sam_hdr_t *h);
void sam_hdr_destroy(sam_hdr_t *h);
int sam_hdr_add_pg(sam_hdr_t *h, const char *name, ...);
int sam_hdr_find_tag_hd(sam_hdr_t *h, const char *key, kstring_t *ks);
int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b);

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src);
bam1_t *bam_dup1(const bam1_t *src);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_del(bam1_t *b, uint8_t *s);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);
hts_pos_t bam_endpos(const bam1_t *b);
char *stringify_argv(int argc, char *argv[]);
#endif
