/* keyorder.c -- see keyorder.h */
#include "keyorder.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { size_t *e; size_t n, cap; } bucket;
struct keyorder {
    char **keys; size_t n, cap;       /* keys in reference iteration order */
    bucket *slot; int level, slots;   /* slot -> indices into keys[] in bucket order */
};

static const double MULT[7] = {3.1415926536, 2.7182818285, 1.6180339887, 1.7320508076, 2.2360679775, 2.6457513111, 3.3166247904};

static int slot_of(const keyorder *k, const char *key)
{
    double sum = 0;
    size_t l = strlen(key);
    for (size_t i = 0; i < l; i++) sum += key[i] * MULT[i % 7];
    return (int)(k->slots * (sum - floor(sum)));
}

static void bucket_push(bucket *b, size_t v)
{
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 4; b->e = realloc(b->e, b->cap * sizeof(size_t)); }
    b->e[b->n++] = v;
}

static void list_push(keyorder *k, char *key)
{
    if (k->n == k->cap) { k->cap = k->cap ? 2 * k->cap : 16; k->keys = realloc(k->keys, k->cap * sizeof(char *)); }
    k->keys[k->n++] = key;
}

static void expand(keyorder *k)
{
    bucket *old = k->slot; int oldslots = k->slots;
    char **oldkeys = k->keys; size_t oldn = k->n;
    k->level++;
    k->slots = (int)pow(4, k->level);
    k->slot = calloc((size_t)k->slots, sizeof(bucket));
    if (oldn == 0) { free(old); return; }
    k->keys = NULL; k->n = k->cap = 0;
    for (int i = 0; i < oldslots; i++)              /* re-insert slot by slot: this is what permutes the key list */
        for (size_t j = 0; j < old[i].n; j++) {
            char *key = oldkeys[old[i].e[j]];
            list_push(k, key);
            bucket_push(&k->slot[slot_of(k, key)], k->n - 1);
        }
    for (int i = 0; i < oldslots; i++) free(old[i].e);
    free(old); free(oldkeys);
}

keyorder *ko_new(void)
{
    keyorder *k = calloc(1, sizeof *k);
    if (k) expand(k);                                /* level 1: 4 slots */
    return k;
}

void ko_free(keyorder *k)
{
    if (!k) return;
    for (size_t i = 0; i < k->n; i++) free(k->keys[i]);
    for (int i = 0; i < k->slots; i++) free(k->slot[i].e);
    free(k->slot); free(k->keys); free(k);
}

long ko_find(const keyorder *k, const char *key)
{
    const bucket *b = &k->slot[slot_of(k, key)];
    for (size_t i = 0; i < b->n; i++) if (!strcmp(k->keys[b->e[i]], key)) return (long)b->e[i];
    return -1;
}

int ko_add(keyorder *k, const char *key)
{
    if (ko_find(k, key) >= 0) return 0;
    char *c = malloc(strlen(key) + 1);
    strcpy(c, key);
    list_push(k, c);
    bucket_push(&k->slot[slot_of(k, key)], k->n - 1);
    if ((float)k->n / (float)k->slots >= 2.0f) expand(k);
    return 1;
}

size_t ko_size(const keyorder *k) { return k->n; }
const char *ko_key(const keyorder *k, size_t i) { return k->keys[i]; }
