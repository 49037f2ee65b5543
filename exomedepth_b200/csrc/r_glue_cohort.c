/* r_glue_cohort.c — ONE `.Call` for a whole cohort: the batched form of the loop R drives around the two reference
 * routines (`new('ExomeDepth')` likelihood step, R/class_definition.R:184-189, then CallCNVs' per-chromosome loop,
 * :354-409, once per sample).  The two-routine drop-in (r_glue.c) costs one host -> device -> host round trip per
 * `.Call` — one per sample for the likelihoods, one per sample x chromosome for the Viterbi; this entry hands the
 * count matrix of the cohort over once and gets every sample's CNV.calls columns back.
 *
 *   edb_cohort_callcnvs(counts, reference, chromosome, start, end, phi, expected,
 *                       transition.probability, expected.CNV.length)
 *     counts      INTSXP  n.bins x n.samples matrix (one column per test sample, as getBamCounts returns them,
 *                         R/countBamInGranges.R:356-369); an R matrix is column-major, i.e. already the
 *                         [sample][bin] layout of the C ABI
 *     reference   INTSXP  n.bins (one aggregate shared by all samples) or n.bins x n.samples (one per sample)
 *     chromosome  INTSXP  n.bins codes of the chromosome factor, bins ALREADY ordered by (chromosome, midpoint) as
 *                         CallCNVs orders them (:323-336) — the R wrapper in INTEGRATION.md does the order()
 *     start, end  INTSXP  n.bins
 *     phi, expected REALSXP n.samples (the per-sample fit; per-bin vectors: n.bins x n.samples matrices)
 *   ->  VECSXP n.samples; element s = REALSXP matrix n.calls x 8, columns
 *         start.p, end.p, type, nexons, BF (before signif), reads.expected (before as.integer), reads.observed,
 *         cor(test, reference) [same value in every row]
 *       i.e. what :371-400 computes; the remaining columns (start, end, chromosome, id, reads.ratio) are lookups
 *       and formatting the R wrapper does on a few hundred rows.
 */
#include <R.h>
#include <Rinternals.h>
#include <math.h>
#include <string.h>

#include "exomedepth_b200.h"

static void raise_if_failed_cohort(int rc, const char *what)
{
    if (rc & (EDB200_ERR_CUDA | EDB200_ERR_ARG | EDB200_ERR_NSTATES))
        error("%s: %s (exomedepth_b200 status %d; this build has no CPU fallback)", what, edb200_last_error(), rc);
}

SEXP edb_cohort_callcnvs(SEXP counts, SEXP reference, SEXP chromosome, SEXP start, SEXP end, SEXP phi, SEXP expected,
                         SEXP transition_probability, SEXP expected_cnv_length)
{
    const int n_bins = length(start);
    const int n_samples = n_bins > 0 ? length(counts) / n_bins : 0;
    const int cap = 4096;
    int n_chains = 0, b, s, k, rc, per_bin;
    int64_t *offsets;
    int32_t *calls, *ncalls;
    double *stats, *cor;
    edb200_cohort_spec spec;
    edb200_cohort *co = NULL;
    edb200_batch batch;
    SEXP out;

    if (n_bins < 1 || n_samples < 1 || length(counts) != n_bins * n_samples) error("counts must be an n.bins x n.samples integer matrix");
    if (length(end) != n_bins || length(chromosome) != n_bins) error("chromosome, start and end must have one entry per bin");
    if (length(reference) != n_bins && length(reference) != n_bins * n_samples) error("reference: n.bins counts, or one column per sample");
    per_bin = length(phi) == n_bins * n_samples && n_bins != 1;
    if (!per_bin && (length(phi) != n_samples || length(expected) != n_samples)) error("phi and expected: one value per sample (or one per bin and sample)");
    if (per_bin && length(expected) != length(phi)) error("phi and expected must have the same shape");

    /* chromosome runs -> chain offsets (the bins arrive ordered, so every chromosome is one run, :354-356) */
    offsets = (int64_t *)R_alloc((size_t)n_bins + 1, sizeof(int64_t));
    offsets[0] = 0;
    for (b = 1; b < n_bins; b++)
        if (INTEGER(chromosome)[b] != INTEGER(chromosome)[b - 1]) offsets[++n_chains] = b;
    offsets[++n_chains] = n_bins;

    memset(&spec, 0, sizeof spec);
    spec.n_bins = n_bins;
    spec.n_chains = n_chains;
    spec.chain_offsets = offsets;
    spec.start = INTEGER(start);
    spec.end = INTEGER(end);
    spec.n_states = 3;                                    /* the reference's model (src/hmm.cpp:37-40) */
    spec.mixture = 1.0;
    spec.transition_probability = REAL(transition_probability)[0];
    spec.expected_cnv_length = REAL(expected_cnv_length)[0];
    rc = edb200_cohort_create(&spec, &co);
    raise_if_failed_cohort(rc, "edb_cohort_callcnvs");

    calls = (int32_t *)R_alloc((size_t)n_samples * cap * 4, sizeof(int32_t));
    ncalls = (int32_t *)R_alloc((size_t)n_samples, sizeof(int32_t));
    stats = (double *)R_alloc((size_t)n_samples * cap * 3, sizeof(double));
    cor = (double *)R_alloc((size_t)n_samples, sizeof(double));
    memset(&batch, 0, sizeof batch);
    batch.n_samples = n_samples;
    batch.observed = INTEGER(counts);
    batch.obs_stride = n_bins;
    batch.reference = INTEGER(reference);
    batch.ref_stride = length(reference) == n_bins ? 0 : n_bins;
    batch.phi = REAL(phi);
    batch.expected = REAL(expected);
    batch.per_bin_stride = per_bin ? n_bins : 0;
    batch.calls = calls;
    batch.ncalls = ncalls;
    batch.call_cap = cap;
    batch.call_stats = stats;
    batch.cor = cor;
    rc = edb200_cohort_run_host(co, &batch, EDB200_EMISSION_AUTO);
    edb200_cohort_destroy(co);
    raise_if_failed_cohort(rc, "edb_cohort_callcnvs");
    if (rc & EDB200_WARN_NAN) Rprintf("ERROR %s %i %s\n", "beta.c", 44, "domain error");      /* src/error.c:45-48, once per call */
    if (rc & EDB200_WARN_CALLCAP) error("edb_cohort_callcnvs: a sample has more than %d calls", cap);

    PROTECT(out = allocVector(VECSXP, n_samples));
    for (s = 0; s < n_samples; s++) {
        const int n = ncalls[s];
        SEXP m = allocMatrix(REALSXP, n, 8);
        double *v = REAL(m);
        for (k = 0; k < n; k++) {
            const int32_t *c = calls + ((size_t)s * cap + k) * 4;
            const double *st = stats + ((size_t)s * cap + k) * 3;
            v[k + 0 * n] = c[0];
            v[k + 1 * n] = c[1];
            v[k + 2 * n] = c[2];
            v[k + 3 * n] = c[3];
            v[k + 4 * n] = 0.43429448190325182765 * st[0];     /* log10(e) * sum(ll[, type] - ll[, normal])  (:393-396) */
            v[k + 5 * n] = st[1];
            v[k + 6 * n] = st[2];
            v[k + 7 * n] = cor[s];
        }
        SET_VECTOR_ELT(out, s, m);
    }
    UNPROTECT(1);
    return out;
}

/* entry to append to CallEntries in r_glue.c:
 *     {"edb_cohort_callcnvs", (DL_FUNC) &edb_cohort_callcnvs, 9},
 */
