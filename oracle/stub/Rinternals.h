/* Minimal stand-in for R's <Rinternals.h>, written for this repo (NOT R source).
 *
 * TEST INFRASTRUCTURE ONLY.  R is not installed on the build/GPU boxes, so the
 * reference's C/C++ (compiled unmodified, in place, from /root/reference/src by
 * oracle/Makefile) and this repo's own .Call glue (exomedepth_b200/csrc/r_glue.c)
 * are both compiled against this header when they are exercised from tests.
 * A "SEXP" here is a pointer to a tiny tagged struct that Python's ctypes can
 * build and read (oracle/sexp.py mirrors the layout).
 *
 * Only the handful of entry points used by src/CNV_estimate.cpp, src/hmm.cpp,
 * src/error.c and src/stream.c of the reference are provided.
 */
#ifndef EDB200_STUB_RINTERNALS_H
#define EDB200_STUB_RINTERNALS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct edb200_stub_sexp {
    int   type;   /* INTSXP / REALSXP / VECSXP                       */
    int   n;      /* number of elements                              */
    int   nrow;   /* matrix rows (0 when not a matrix)               */
    int   ncol;   /* matrix cols (0 when not a matrix)               */
    void *data;   /* int[n], double[n] or SEXP[n]                    */
} *SEXP;

#define INTSXP  13
#define REALSXP 14
#define VECSXP  19

typedef int R_xlen_t;

double *REAL(SEXP x);
int    *INTEGER(SEXP x);
int     length(SEXP x);
SEXP    allocVector(int type, int n);
SEXP    allocMatrix(int type, int nrow, int ncol);
SEXP    SET_VECTOR_ELT(SEXP list, int i, SEXP v);
SEXP    VECTOR_ELT(SEXP list, int i);
/* transient storage R reclaims when .Call returns; the stand-in reclaims it in edb200_stub_free */
char   *R_alloc(size_t n, int size);
void    Rprintf(const char *fmt, ...);
void    REprintf(const char *fmt, ...);
void    Rf_error(const char *fmt, ...);
void    Rf_warning(const char *fmt, ...);
/* test helper: free a SEXP tree made by allocVector/allocMatrix */
void    edb200_stub_free(SEXP x);
/* test helper: number of lines printed through Rprintf since last reset */
int     edb200_stub_rprintf_count(int reset);
/* test helper: silence (1) or restore (0) Rprintf output to stderr */
void    edb200_stub_rprintf_quiet(int quiet);

#define PROTECT(x)   (x)
#define UNPROTECT(n) ((void)(n))
#define R_NilValue   ((SEXP)0)
#define error        Rf_error
#define warning      Rf_warning

#ifdef __cplusplus
}
#endif
#endif
