/* r_glue_ext.c — `.Call` routines for the rows AROUND the hot path, on top of the C ABI.
 *
 * These have no counterpart in the reference's registration table (src/ExomeDepth_init.c:14-18): they replace
 * third-party R code the package calls — aod::betabin / aod::fitted in `new('ExomeDepth')`
 * (R/class_definition.R:118-119, 168) and the per-candidate cor() of select.reference.set
 * (R/optimize_reference_set.R:100).  A maintainer who wants them adds the two entries below to CallEntries in
 * r_glue.c and the R-side calls shown in INTEGRATION.md; r_glue.c alone stays the exact drop-in.
 *
 *   edb_betabin_fit(test, reference)                    INTSXP n, INTSXP n  ->  REALSXP 4: phi, expected, loglik, info
 *   edb_refset_correlations(test, reference.counts,     INTSXP n, INTSXP n x m (R matrix: one column per candidate),
 *                           bin.length, selected)       REALSXP n (or length 0: all 1), INTSXP k (1-based, as which())
 *                                                       ->  REALSXP m: the correlations of :100
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "exomedepth_b200.h"

static void raise_if_failed_ext(int rc, const char *what)
{
    if (rc & (EDB200_ERR_CUDA | EDB200_ERR_ARG | EDB200_ERR_NSTATES))
        error("%s: %s (exomedepth_b200 status %d; this build has no CPU fallback)", what, edb200_last_error(), rc);
}

SEXP edb_betabin_fit(SEXP test, SEXP reference)
{
    const int n = length(test);
    double mu = 0, phi = 0, ll = 0;
    int32_t info = 0;
    SEXP out;
    int rc;
    if (length(reference) != n) error("Length of test and numeric must match");          /* R/class_definition.R:92 */
    rc = edb200_betabin_fit(INTEGER(test), n, INTEGER(reference), 0, 1, n, &mu, &phi, &ll, &info);
    raise_if_failed_ext(rc, "edb_betabin_fit");
    PROTECT(out = allocVector(REALSXP, 4));
    REAL(out)[0] = phi;
    REAL(out)[1] = mu;
    REAL(out)[2] = ll;
    REAL(out)[3] = info;
    UNPROTECT(1);
    return out;
}

SEXP edb_refset_correlations(SEXP test, SEXP reference_counts, SEXP bin_length, SEXP selected)
{
    const int n = length(test), k = length(selected);
    const int m = n > 0 ? length(reference_counts) / n : 0;
    int32_t *stacked, *sel0;
    double *row;
    SEXP out;
    int rc, i;
    if (n < 1 || m < 1 || length(reference_counts) != n * m)
        error("The number of rows of the reference matrix must match the length of the test count data");   /* optimize_reference_set.R:64 */
    if (length(bin_length) != 0 && length(bin_length) != n) error("bin.length must have one entry per bin");
    /* an R integer matrix is column-major: candidate j is the contiguous block [j*n, (j+1)*n) — already the
       [sample][bin] layout of the C ABI; the test sample goes in front as row 0 */
    stacked = (int32_t *)R_alloc((size_t)(m + 1) * n, sizeof(int32_t));
    memcpy(stacked, INTEGER(test), (size_t)n * sizeof(int32_t));
    memcpy(stacked + n, INTEGER(reference_counts), (size_t)n * m * sizeof(int32_t));
    sel0 = (int32_t *)R_alloc((size_t)(k > 0 ? k : 1), sizeof(int32_t));
    for (i = 0; i < k; i++) sel0[i] = INTEGER(selected)[i] - 1;
    row = (double *)R_alloc((size_t)(m + 1), sizeof(double));
    rc = edb200_refset_correlations(stacked, n, m + 1, length(bin_length) ? REAL(bin_length) : NULL, sel0, k, 0, 1, row);
    raise_if_failed_ext(rc, "edb_refset_correlations");
    PROTECT(out = allocVector(REALSXP, m));
    for (i = 0; i < m; i++) REAL(out)[i] = row[i + 1];
    UNPROTECT(1);
    return out;
}

/* entries to append to CallEntries in r_glue.c:
 *     {"edb_betabin_fit",         (DL_FUNC) &edb_betabin_fit,         2},
 *     {"edb_refset_correlations", (DL_FUNC) &edb_refset_correlations, 4},
 */
