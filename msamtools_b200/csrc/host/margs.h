/* margs.h -- command-line tables for the drop-in CLI.  The reference parses its options with
 * argtable2 (a system library, configure.ac:77-88, not installed here); this is a from-scratch
 * parser with the same table-driven shape and the same grammar (argtable2 fronts getopt_long):
 * bundled short flags, "-l 80" / "-l80", "--opt v" / "--opt=v", negative numbers accepted as
 * option values, positional arguments go to the arg_file entry. */
#ifndef MSG_MARGS_H
#define MSG_MARGS_H
#include <stdio.h>

enum { ARG_TERMINATOR = 1, ARG_HASVALUE = 2 };
typedef struct arg_hdr {
    char flag; int kind;                       /* kind: 'l' lit, 'i' int, 's' str, 'f' file, 'e' end */
    const char *shortopts, *longopts, *datatype, *glossary;
    int mincount, maxcount;
} arg_hdr;
struct arg_lit  { arg_hdr hdr; int count; };
struct arg_int  { arg_hdr hdr; int count; int *ival; };
struct arg_str  { arg_hdr hdr; int count; const char **sval; };
struct arg_file { arg_hdr hdr; int count; const char **filename, **basename, **extension; };
struct arg_end  { arg_hdr hdr; int count; int *error; void **parent; const char **argval; };

struct arg_lit  *arg_lit0(const char *s, const char *l, const char *glossary);
struct arg_int  *arg_int0(const char *s, const char *l, const char *datatype, const char *glossary);
struct arg_str  *arg_str0(const char *s, const char *l, const char *datatype, const char *glossary);
struct arg_str  *arg_str1(const char *s, const char *l, const char *datatype, const char *glossary);
struct arg_file *arg_filen(const char *s, const char *l, const char *datatype, int mincount, int maxcount, const char *glossary);
struct arg_end  *arg_end(int maxerrors);
int  arg_nullcheck(void **argtable);
int  arg_parse(int argc, char **argv, void **argtable);
void arg_print_errors(FILE *fp, struct arg_end *end, const char *progname);
void arg_print_syntax(FILE *fp, void **argtable, const char *suffix);
void arg_print_glossary(FILE *fp, void **argtable, const char *format);
void arg_freetable(void **argtable, size_t n);
#endif
