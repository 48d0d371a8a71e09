/*
 * keyorder.h -- insertion-ordered string set whose iteration order reproduces the reference's
 * `zoeKeysOfHash` (zoeTools.c:218-279,335-372): with --genome the profile's feature rows come out
 * in that order (msam_profile.c:791-798,837-843), so a drop-in has to match it.
 *
 * Order rule: keys are listed in insertion order, except that every time the table grows
 * (4^level slots, grown when keys/slots >= 2) the list is rebuilt by walking the OLD table slot
 * by slot, bucket order within a slot.  The slot of a key is floor(slots * frac(sum_i key[i] *
 * M[i % 7])) with the seven multipliers pi, e, phi, sqrt 3, sqrt 5, sqrt 7, sqrt 11.
 */
#ifndef MSG_KEYORDER_H
#define MSG_KEYORDER_H
#include <stddef.h>
typedef struct keyorder keyorder;
keyorder *ko_new(void);
void ko_free(keyorder *k);
/* insert if absent; returns 1 if new, 0 if already present */
int ko_add(keyorder *k, const char *key);
size_t ko_size(const keyorder *k);
const char *ko_key(const keyorder *k, size_t i);     /* i-th key in reference order */
long ko_find(const keyorder *k, const char *key);    /* position in reference order or -1 */
#endif
