/* Minimal stand-in for htslib/kstring.h (htslib 1.24 is not vendored by the reference and not
 * installed here).  TEST INFRASTRUCTURE: only used to compile the reference's own sources into
 * oracle/_ref/.  kstrtok keeps htslib's semantics: empty tokens are returned, not skipped. */
#ifndef SHIM_KSTRING_H
#define SHIM_KSTRING_H
#include <stddef.h>
#include <stdint.h>
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
typedef struct ks_tokaux_t { uint64_t tab[4]; int sep, finished; const char *p; } ks_tokaux_t;
char *kstrtok(const char *str, const char *sep, ks_tokaux_t *aux);
#endif
