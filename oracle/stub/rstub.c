/* Implementation of the stand-in R API declared in Rinternals.h (this directory).
 * Test infrastructure only; written for this repo. */
#include "Rinternals.h"
#include "R_ext/Rdynload.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

static int g_rprintf_lines = 0;
static int g_rprintf_quiet = 0;

double *REAL(SEXP x)    { return (double *)x->data; }
int    *INTEGER(SEXP x) { return (int *)x->data; }
int     length(SEXP x)  { return x ? x->n : 0; }

SEXP allocVector(int type, int n)
{
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    size_t elt = type == INTSXP ? sizeof(int) : type == REALSXP ? sizeof(double) : sizeof(SEXP);
    s->type = type;
    s->n = n;
    s->data = calloc(n > 0 ? (size_t)n : 1, elt);
    return s;
}

SEXP allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = allocVector(type, nrow * ncol);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}

SEXP SET_VECTOR_ELT(SEXP list, int i, SEXP v) { ((SEXP *)list->data)[i] = v; return v; }
SEXP VECTOR_ELT(SEXP list, int i)             { return ((SEXP *)list->data)[i]; }

struct transient { struct transient *next; };
static struct transient *g_transient = NULL;

char *R_alloc(size_t n, int size)
{
    struct transient *t = (struct transient *)malloc(sizeof(*t) + 16 + n * (size_t)size);
    t->next = g_transient;
    g_transient = t;
    return (char *)t + 16;
}

static void free_transient(void)
{
    while (g_transient) {
        struct transient *t = g_transient;
        g_transient = t->next;
        free(t);
    }
}

void edb200_stub_free(SEXP x)
{
    free_transient();
    if (!x) return;
    if (x->type == VECSXP)
        for (int i = 0; i < x->n; i++) edb200_stub_free(((SEXP *)x->data)[i]);
    free(x->data);
    free(x);
}

void Rprintf(const char *fmt, ...)
{
    va_list ap;
    g_rprintf_lines++;
    if (g_rprintf_quiet) return;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

void REprintf(const char *fmt, ...)
{
    va_list ap;
    if (g_rprintf_quiet) return;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

void Rf_warning(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fputs("Warning: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}

void Rf_error(const char *fmt, ...)
{
    /* real R longjmps back to the top level; the stand-in records and aborts */
    va_list ap;
    va_start(ap, fmt);
    fputs("Error: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
    abort();
}

int edb200_stub_rprintf_count(int reset)
{
    int n = g_rprintf_lines;
    if (reset) g_rprintf_lines = 0;
    return n;
}

void edb200_stub_rprintf_quiet(int quiet) { g_rprintf_quiet = quiet; }

int R_registerRoutines(DllInfo *dll, const void *c_entries, const R_CallMethodDef *call_entries,
                       const void *fortran_entries, const void *external_entries)
{
    (void)c_entries; (void)fortran_entries; (void)external_entries;
    if (dll) dll->call_entries = call_entries;
    return 1;
}

int R_useDynamicSymbols(DllInfo *dll, Rboolean value)
{
    if (dll) dll->use_dynamic_symbols = value;
    return 1;
}
