/* r_glue.c — the R `.Call` boundary of ExomeDepth on top of the B200 C ABI (include/exomedepth_b200.h).
 *
 * Drop-in for the three symbols R binds (paths relative to /root/reference):
 *   get_loglike_matrix(phi, expected, total, observed, mixture)            src/CNV_estimate.cpp:52-85
 *   C_hmm(nstates, nobs, transitions, probabilities, positions, expLength) src/hmm.cpp:18-167
 *   R_init_ExomeDepth(DllInfo*)                                            src/ExomeDepth_init.c:11-24
 * called from R/class_definition.R:184-189 and R/tools.R:97.  Same argument order, SEXP types, return
 * shapes (REALSXP n x 3 column-major; VECSXP[2] = {REALSXP[nobs], REALSXP[ncalls x 4]}), PROTECT balance
 * and messages.  The arithmetic is done by the GPU library; there is NO CPU fallback: without a usable
 * sm_100 device the routines raise an R error carrying edb200_last_error().
 *
 * Built as ExomeDepth.so against a real R (`R CMD SHLIB r_glue.c -lexomedepth_b200`), or against the
 * stand-in R API of oracle/stub/ for the tests (R is not installed on the build / GPU boxes).
 */
#include <R.h>
#include <Rinternals.h>
#include <stdlib.h>
#include <R_ext/Rdynload.h>

#include "exomedepth_b200.h"

static void raise_if_failed(int rc, const char *what)
{
    if (rc & (EDB200_ERR_CUDA | EDB200_ERR_ARG | EDB200_ERR_NSTATES))
        error("%s: %s (exomedepth_b200 status %d; this build has no CPU fallback)", what, edb200_last_error(), rc);
}

/* src/error.c:45-48: every failing GSL call prints its file, line and reason, and evaluation continues (abort is commented
 * out, :51).  The device logs the failing cells with their error sites; the text is the reference's, cell by cell. */
static void report_domain_errors(int rc)
{
    if (rc & EDB200_WARN_NAN) {
        char buf[4096];
        int64_t next = 0, shown = 0;
        const int64_t raised = edb200_gsl_error_log(buf, sizeof buf, 0, &next);
        while (next > shown) {
            Rprintf("%s", buf);
            shown = next;
            edb200_gsl_error_log(buf, sizeof buf, shown, &next);
        }
        if (raised > shown) Rprintf("... and %ld more cells with GSL errors (not listed)\n", (long)(raised - shown));
        if (raised == 0) {      /* (a status without a log: should not happen) */
            Rprintf("ERROR %s %i %s\n", "beta.c", 44, "domain error");
            Rprintf("Default GSL error handler invoked.\n");
        }
    }
}

SEXP get_loglike_matrix(SEXP phi_a, SEXP expected_a, SEXP total_a, SEXP observed_a, SEXP mixture_a)
{
    const int     n        = length(total_a);                    /* src/CNV_estimate.cpp:57 */
    const double *phi      = REAL(phi_a);
    const double *expected = REAL(expected_a);
    const int    *total    = INTEGER(total_a);
    const int    *observed = INTEGER(observed_a);
    const double  mixture  = *REAL(mixture_a);
    SEXP rans;
    int rc;

    if (mixture != 1)                                            /* src/CNV_estimate.cpp:61 */
        Rprintf("As a warning (this could be normal), the mixture coefficient is %f\n", mixture);

    PROTECT(rans = allocMatrix(REALSXP, n, 3));                  /* src/CNV_estimate.cpp:69 */
    rc = edb200_get_loglike_matrix(phi, expected, total, observed, mixture, (int64_t)n, REAL(rans));
    UNPROTECT(1);
    raise_if_failed(rc, "get_loglike_matrix");
    report_domain_errors(rc);
    return rans;
}

SEXP C_hmm(SEXP nstates, SEXP nobs, SEXP transitions, SEXP probabilities, SEXP positions, SEXP expectedLength)
{
    const int nstates_c = *INTEGER(nstates);
    const int nobs_c    = *INTEGER(nobs);
    SEXP myList = NULL, final, calls_R;
    int32_t *path, *calls, ncalls = 0;
    int rc, i, j;

    if (nstates_c != 3) {                                        /* src/hmm.cpp:37-40 */
        Rprintf("ERROR: The code must assume 3 states\n");
        return myList;
    }
    /* every observation can close at most one call (src/hmm.cpp:110-121) */
    path  = (int32_t *)R_alloc((size_t)(nobs_c > 0 ? nobs_c : 1), sizeof(int32_t));
    calls = (int32_t *)R_alloc((size_t)(nobs_c > 0 ? nobs_c : 1) * 4, sizeof(int32_t));
    rc = edb200_hmm(nstates_c, nobs_c, REAL(transitions), REAL(probabilities), INTEGER(positions),
                    *REAL(expectedLength), path, calls, nobs_c > 0 ? nobs_c : 1, &ncalls);
    raise_if_failed(rc, "C_hmm");

    PROTECT(myList = allocVector(VECSXP, 2));                    /* src/hmm.cpp:133-135 */
    PROTECT(final = allocVector(REALSXP, nobs_c));
    PROTECT(calls_R = allocMatrix(REALSXP, ncalls, 4));
    for (i = 0; i < nobs_c; i++) REAL(final)[i] = path[i];       /* src/hmm.cpp:139-141 */
    SET_VECTOR_ELT(myList, 0, final);
    for (i = 0; i < ncalls; i++)                                 /* src/hmm.cpp:144-149: column-major ncalls x 4 */
        for (j = 0; j < 4; j++) REAL(calls_R)[ncalls * j + i] = calls[4 * i + j];
    SET_VECTOR_ELT(myList, 1, calls_R);
    UNPROTECT(3);
    return myList;
}

static const R_CallMethodDef CallEntries[] = {                   /* src/ExomeDepth_init.c:14-18 */
    {"C_hmm",              (DL_FUNC) &C_hmm,              6},
    {"get_loglike_matrix", (DL_FUNC) &get_loglike_matrix, 5},
    {NULL, NULL, 0}
};

void R_init_ExomeDepth(DllInfo *dll)                             /* src/ExomeDepth_init.c:20-24 */
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
