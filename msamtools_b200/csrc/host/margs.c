/*
 * margs.c -- table-driven option parser of the drop-in CLI (see margs.h).  Grammar: short flags may
 * be bundled ("-bu"), a short option's value may be attached or separate ("-l80", "-l 80"), long
 * options take "--opt=value" or "--opt value", a value may be a negative number ("--ppt -980"),
 * "--" ends options, everything else is positional and lands in the arg_file entry.
 */
#include "margs.h"
#include <stdlib.h>
#include <string.h>

enum { EMINCOUNT = 1, EMAXCOUNT, EBADINT, EMISSINGVAL, EUNKNOWN_SHORT, EUNKNOWN_LONG, EEXTRA };

static void init_hdr(arg_hdr *h, int kind, const char *s, const char *l, const char *dt, const char *g, int mn, int mx)
{
    h->flag = (kind != 'l' && kind != 'e') ? ARG_HASVALUE : 0; h->kind = kind;
    h->shortopts = s; h->longopts = l; h->datatype = dt; h->glossary = g; h->mincount = mn; h->maxcount = mx;
}

struct arg_lit *arg_lit0(const char *s, const char *l, const char *g)
{
    struct arg_lit *a = calloc(1, sizeof *a);
    if (a) init_hdr(&a->hdr, 'l', s, l, NULL, g, 0, 1);
    return a;
}
struct arg_int *arg_int0(const char *s, const char *l, const char *dt, const char *g)
{
    struct arg_int *a = calloc(1, sizeof *a);
    if (!a) return NULL;
    init_hdr(&a->hdr, 'i', s, l, dt ? dt : "<int>", g, 0, 1);
    a->ival = calloc(4, sizeof(int));
    return a;
}
static struct arg_str *arg_strn(const char *s, const char *l, const char *dt, const char *g, int mn)
{
    struct arg_str *a = calloc(1, sizeof *a);
    if (!a) return NULL;
    init_hdr(&a->hdr, 's', s, l, dt ? dt : "<string>", g, mn, 1);
    a->sval = calloc(4, sizeof(char *));
    a->sval[0] = "";
    return a;
}
struct arg_str *arg_str0(const char *s, const char *l, const char *dt, const char *g) { return arg_strn(s, l, dt, g, 0); }
struct arg_str *arg_str1(const char *s, const char *l, const char *dt, const char *g) { return arg_strn(s, l, dt, g, 1); }
struct arg_file *arg_filen(const char *s, const char *l, const char *dt, int mn, int mx, const char *g)
{
    struct arg_file *a = calloc(1, sizeof *a);
    if (!a) return NULL;
    init_hdr(&a->hdr, 'f', s, l, dt ? dt : "<file>", g, mn, mx);
    int cap = mx > 0 ? mx + 8 : 8;
    a->filename = calloc((size_t)cap, sizeof(char *)); a->basename = calloc((size_t)cap, sizeof(char *)); a->extension = calloc((size_t)cap, sizeof(char *));
    a->filename[0] = a->basename[0] = a->extension[0] = "";
    return a;
}
struct arg_end *arg_end(int maxerrors)
{
    struct arg_end *a = calloc(1, sizeof *a);
    if (!a) return NULL;
    init_hdr(&a->hdr, 'e', NULL, NULL, NULL, NULL, 1, maxerrors);
    a->hdr.flag = ARG_TERMINATOR;
    a->error = calloc((size_t)maxerrors, sizeof(int)); a->parent = calloc((size_t)maxerrors, sizeof(void *)); a->argval = calloc((size_t)maxerrors, sizeof(char *));
    return a;
}

static int table_len(void **t) { int n = 0; while (!(((arg_hdr *)t[n])->flag & ARG_TERMINATOR)) n++; return n; }

int arg_nullcheck(void **t)
{
    for (int i = 0;; i++) { if (!t[i]) return 1; if (((arg_hdr *)t[i])->flag & ARG_TERMINATOR) return 0; }
}

static void add_err(struct arg_end *e, int code, void *parent, const char *val)
{
    if (e->count < e->hdr.maxcount) { e->error[e->count] = code; e->parent[e->count] = parent; e->argval[e->count] = val; }
    e->count++;
}

static int long_matches(const char *longopts, const char *name, size_t n)
{   /* longopts may be a comma separated list */
    const char *p = longopts;
    while (p && *p) {
        const char *c = strchr(p, ','); size_t l = c ? (size_t)(c - p) : strlen(p);
        if (l == n && !strncmp(p, name, n)) return 1;
        p = c ? c + 1 : NULL;
    }
    return 0;
}

static void take_value(void *ent, const char *val, struct arg_end *e)
{
    arg_hdr *h = ent;
    if (h->kind == 'i') {
        struct arg_int *a = ent;
        char *ep; long v = strtol(val, &ep, 0);
        if (*val == 0 || *ep != 0) { add_err(e, EBADINT, ent, val); return; }
        if (a->count >= h->maxcount) { add_err(e, EMAXCOUNT, ent, val); return; }
        a->ival[a->count++] = (int)v;
    } else if (h->kind == 's') {
        struct arg_str *a = ent;
        if (a->count >= h->maxcount) { add_err(e, EMAXCOUNT, ent, val); return; }
        a->sval[a->count++] = val;
    } else if (h->kind == 'f') {
        struct arg_file *a = ent;
        /* like argtable2, keep counting beyond maxcount only up to the allocated slots; the reference checks count > 1 itself */
        if (a->count >= h->maxcount + 7) { add_err(e, EMAXCOUNT, ent, val); return; }
        a->filename[a->count] = val;
        const char *b = strrchr(val, '/'); a->basename[a->count] = b ? b + 1 : val;
        const char *x = strrchr(a->basename[a->count], '.'); a->extension[a->count] = x ? x : "";
        a->count++;
    }
}

int arg_parse(int argc, char **argv, void **t)
{
    const int n = table_len(t);
    struct arg_end *e = t[n];
    e->count = 0;
    void *positional = NULL;
    for (int i = 0; i < n; i++) { arg_hdr *h = t[i]; if (!h->shortopts && !h->longopts && h->kind == 'f') positional = t[i]; }
    int only_positional = 0;
    for (int k = 1; k < argc; k++) {
        const char *a = argv[k];
        if (!only_positional && a[0] == '-' && a[1] == '-' && a[2] == 0) { only_positional = 1; continue; }
        if (!only_positional && a[0] == '-' && a[1] == '-') {
            const char *name = a + 2, *eq = strchr(name, '=');
            size_t nl = eq ? (size_t)(eq - name) : strlen(name);
            void *ent = NULL;
            for (int i = 0; i < n; i++) { arg_hdr *h = t[i]; if (h->longopts && long_matches(h->longopts, name, nl)) { ent = t[i]; break; } }
            if (!ent) { add_err(e, EUNKNOWN_LONG, NULL, a); continue; }
            arg_hdr *h = ent;
            if (h->kind == 'l') { struct arg_lit *l = ent; if (l->count < h->maxcount) l->count++; else add_err(e, EMAXCOUNT, ent, a); continue; }
            const char *val = eq ? eq + 1 : (k + 1 < argc ? argv[++k] : NULL);
            if (!val) { add_err(e, EMISSINGVAL, ent, a); continue; }
            take_value(ent, val, e);
            continue;
        }
        if (!only_positional && a[0] == '-' && a[1] != 0) {
            for (const char *c = a + 1; *c; c++) {
                void *ent = NULL;
                for (int i = 0; i < n; i++) { arg_hdr *h = t[i]; if (h->shortopts && strchr(h->shortopts, *c)) { ent = t[i]; break; } }
                if (!ent) { add_err(e, EUNKNOWN_SHORT, NULL, a); break; }
                arg_hdr *h = ent;
                if (h->kind == 'l') { struct arg_lit *l = ent; if (l->count < h->maxcount) l->count++; else add_err(e, EMAXCOUNT, ent, a); continue; }
                const char *val = c[1] ? c + 1 : (k + 1 < argc ? argv[++k] : NULL);
                if (!val) add_err(e, EMISSINGVAL, ent, a); else take_value(ent, val, e);
                break;
            }
            continue;
        }
        if (positional) take_value(positional, a, e); else add_err(e, EEXTRA, NULL, a);
    }
    for (int i = 0; i < n; i++) {
        arg_hdr *h = t[i];
        int count = h->kind == 'l' ? ((struct arg_lit *)t[i])->count : h->kind == 'i' ? ((struct arg_int *)t[i])->count :
                    h->kind == 's' ? ((struct arg_str *)t[i])->count : ((struct arg_file *)t[i])->count;
        if (count < h->mincount) add_err(e, EMINCOUNT, t[i], NULL);
    }
    return e->count;
}

static void opt_name(const arg_hdr *h, char *buf, size_t n)
{
    if (h->shortopts) snprintf(buf, n, "-%c", h->shortopts[0]);
    else if (h->longopts) { const char *c = strchr(h->longopts, ','); snprintf(buf, n, "--%.*s", (int)(c ? c - h->longopts : (long)strlen(h->longopts)), h->longopts); }
    else snprintf(buf, n, "%s", h->datatype ? h->datatype : "");
}

void arg_print_errors(FILE *fp, struct arg_end *e, const char *progname)
{
    for (int i = 0; i < e->count && i < e->hdr.maxcount; i++) {
        char nm[128] = ""; const arg_hdr *h = e->parent[i];
        if (h) opt_name(h, nm, sizeof nm);
        const char *dt = h && h->datatype && (h->shortopts || h->longopts) ? h->datatype : "";
        switch (e->error[i]) {
        case EMINCOUNT: fprintf(fp, "%s: missing option %s %s\n", progname, nm, dt); break;
        case EMAXCOUNT: fprintf(fp, "%s: excess option %s %s\n", progname, nm, e->argval[i] ? e->argval[i] : ""); break;
        case EBADINT: fprintf(fp, "%s: invalid argument \"%s\" to option %s %s\n", progname, e->argval[i], nm, dt); break;
        case EMISSINGVAL: fprintf(fp, "%s: option \"%s\" requires an argument\n", progname, e->argval[i]); break;
        case EUNKNOWN_SHORT: case EUNKNOWN_LONG: fprintf(fp, "%s: invalid option \"%s\"\n", progname, e->argval[i]); break;
        default: fprintf(fp, "%s: unexpected argument \"%s\"\n", progname, e->argval[i]); break;
        }
    }
}

void arg_print_syntax(FILE *fp, void **t, const char *suffix)
{
    const int n = table_len(t);
    /* bundled flags first: [-bhS] */
    char flags[64]; int nf = 0;
    for (int i = 0; i < n; i++) { arg_hdr *h = t[i]; if (h->kind == 'l' && h->shortopts && h->mincount == 0) flags[nf++] = h->shortopts[0]; }
    flags[nf] = 0;
    if (nf) fprintf(fp, " [-%s]", flags);
    for (int i = 0; i < n; i++) {
        arg_hdr *h = t[i];
        if (h->kind == 'l' && h->shortopts) continue;
        char nm[128]; opt_name(h, nm, sizeof nm);
        const char *open = h->mincount == 0 ? "[" : "", *close = h->mincount == 0 ? "]" : "";
        if (h->kind == 'l') fprintf(fp, " %s%s%s", open, nm, close);
        else if (h->shortopts || h->longopts) fprintf(fp, " %s%s%s%s%s", open, nm, h->shortopts ? " " : "=", h->datatype, close);
        else fprintf(fp, " %s%s%s", open, h->datatype, close);
    }
    fputs(suffix ? suffix : "", fp);
}

void arg_print_glossary(FILE *fp, void **t, const char *format)
{
    const int n = table_len(t);
    for (int i = 0; i < n; i++) {
        arg_hdr *h = t[i];
        if (!h->glossary) continue;
        char syn[256] = ""; size_t l = 0;
        if (h->shortopts) l += (size_t)snprintf(syn + l, sizeof syn - l, "-%c", h->shortopts[0]);
        if (h->longopts) l += (size_t)snprintf(syn + l, sizeof syn - l, "%s--%s", l ? ", " : "", h->longopts);
        if (h->kind != 'l' && h->datatype) l += (size_t)snprintf(syn + l, sizeof syn - l, "%s%s", (h->shortopts || h->longopts) ? (h->longopts ? "=" : " ") : "", h->datatype);
        fprintf(fp, format, syn, h->glossary);
    }
}

void arg_freetable(void **t, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        if (!t[i]) continue;
        arg_hdr *h = t[i];
        if (h->kind == 'i') free(((struct arg_int *)t[i])->ival);
        else if (h->kind == 's') free(((struct arg_str *)t[i])->sval);
        else if (h->kind == 'f') { struct arg_file *a = t[i]; free(a->filename); free(a->basename); free(a->extension); }
        else if (h->kind == 'e') { struct arg_end *a = t[i]; free(a->error); free(a->parent); free(a->argval); }
        free(t[i]); t[i] = NULL;
    }
}
