/* exomedepth_b200.h — C ABI of the B200-native CNV-calling core for ExomeDepth.
 *
 * This is the drop-in boundary for ONE hot path of the reference (paths relative to /root/reference):
 *   - the per-bin beta-binomial emission log-likelihood   src/CNV_estimate.cpp:52-85 (get_loglike_matrix)
 *   - the HMM Viterbi sweep + traceback + segment summary  src/hmm.cpp:18-167        (C_hmm)
 * as registered for .Call in src/ExomeDepth_init.c:14-24 and called from R/class_definition.R:184-189
 * and R/tools.R:97.  Plain pointers and sizes only; no R, torch or CUDA types.
 *
 * Every entry point runs on the selected B200 (sm_100a).  There is NO CPU fallback: without a usable
 * device the calls return EDB200_ERR_CUDA and edb200_last_error() says why.
 *
 * Threading: the host-pointer entry points serialise on an internal mutex.  The device-pointer entry points only
 * enqueue work; they share the context's helper streams and scratch buffers, so calls into one context must not be
 * issued concurrently from several host threads (one process per GPU is the intended deployment).
 *
 * Return value of every int function: 0 on success, otherwise a bit mask of the EDB200_* codes.
 * EDB200_WARN_NAN alone is not a failure: like the reference (src/error.c:35-52, error.h:9) a GSL-style
 * domain error yields NaN in the affected cells, a message, and the computation continues.
 */
#ifndef EXOMEDEPTH_B200_H
#define EXOMEDEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define EDB200_API __attribute__((visibility("default")))
#else
#define EDB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EDB200_OK            0
#define EDB200_WARN_NAN      1   /* domain error(s): NaN cells produced, as src/beta.c:43-45 / error.h:9 */
#define EDB200_ERR_NSTATES   2   /* unsupported number of states (cf. src/hmm.cpp:37-40)                  */
#define EDB200_ERR_CUDA      4   /* no device / CUDA failure                                             */
#define EDB200_ERR_ARG       8   /* bad argument                                                         */
#define EDB200_WARN_CALLCAP 16   /* more CNV calls than the caller's buffer holds; ncalls has the count  */

#define EDB200_MAX_STATES    7

/* ---- lifetime ---------------------------------------------------------------------------------- */
/* Select the device and create the context (idempotent). device < 0: use LOCAL_RANK or 0. */
EDB200_API int         edb200_init(int device);
EDB200_API void        edb200_shutdown(void);
EDB200_API const char *edb200_last_error(void);
/* "NVIDIA B200 sm_100 148 SMs" style description of the selected device */
EDB200_API int         edb200_device_info(char *buf, int buflen, int *n_sms, int *cc_major, int *cc_minor);
/* count of kernels launched by this library since the last reset (bench.py's gpu_launches) */
EDB200_API int64_t     edb200_launch_count(int reset);
/* per-kernel device times: while enabled, every kernel launch of this library is bracketed by CUDA events on its
 * stream; edb200_profile_read synchronises and writes "name:launches:total_ms:longest_ms;..." (bench.py's roofline uses it) */
EDB200_API int         edb200_profile(int enable);   /* 0 off, 1 on, 2 on + print start / duration of every launch to stderr on read */
EDB200_API int         edb200_profile_read(char *buf, int buflen);
/* page-locked host buffers for the host-pointer entry points (optional, speeds up the copies) */
EDB200_API void       *edb200_host_alloc(size_t bytes);
EDB200_API void        edb200_host_free(void *p);
/* Encoders of the count-matrix ingestion layouts of edb200_batch (observed16 / observed12, below) — what a loader runs once per
 * cohort over the integer matrix getBamCounts produced (R/countBamInGranges.R:356-369).  Host code on the host's threads; no
 * GPU involved.  counts: int32 [n_samples][stride]; out16: uint16 [n_samples][out_stride] (elements), out12: rows out_stride
 * BYTES apart, >= 3 * ceil(n_bins / 2).  Counts at or beyond the layout's sentinel (65535 / 4095) go to the overflow list,
 * sorted by flat index sample * n_bins + bin; the first overflow_cap entries are written.  Returns the number of entries the
 * matrix has (call again with larger lists if it exceeds overflow_cap), -1 for a bad argument or a negative count. */
EDB200_API int64_t     edb200_pack_counts16(const int32_t *counts, int64_t stride, int32_t n_samples, int64_t n_bins, uint16_t *out16,
                                            int64_t out_stride, int64_t *overflow_index, int32_t *overflow_value, int64_t overflow_cap);
EDB200_API int64_t     edb200_pack_counts12(const int32_t *counts, int64_t stride, int32_t n_samples, int64_t n_bins, uint8_t *out12,
                                            int64_t out_stride, int64_t *overflow_index, int32_t *overflow_value, int64_t overflow_cap);

/* ---- the two .Call routines, argument for argument --------------------------------------------- */

/* Replaces get_loglike_matrix(phi, expected, total, observed, mixture)   src/CNV_estimate.cpp:52-85.
 * phi, expected: double[n]; total, observed: int32[n]; ll_out: double[n*3] column-major
 * (ll_out[c + n*s], s = 0 deletion, 1 normal, 2 duplication).  Host pointers. */
EDB200_API int edb200_get_loglike_matrix(const double *phi, const double *expected, const int32_t *total,
                              const int32_t *observed, double mixture, int64_t n, double *ll_out);

/* gsl_sf_lnbeta(x, y) (src/beta.c:161-164: the only symbol of the vendored GSL that the hot path calls,
 * src/CNV_estimate.cpp:33,49) evaluated element-wise by the device's restatement of that chain — the code the
 * emission kernels fall back to for non-positive shape parameters.  NaN + EDB200_WARN_NAN where the reference
 * prints a domain error.  Host pointers; exists so that the special functions can be pinned on their own. */
EDB200_API int edb200_lnbeta(const double *x, const double *y, int64_t n, double *out);

/* S-state generalisation of the same routine (extension; n_states = 3 with odds = NULL or
 * {1-mix/2, 1, 1+mix/2} is exactly the call above).  odds: double[n_states] multiplying the odds of the
 * expected proportion per state; ll_out: double[n*n_states] column-major. */
EDB200_API int edb200_emission(const double *phi, const double *expected, const int32_t *total,
                    const int32_t *observed, int64_t n, int32_t n_states, const double *odds, double *ll_out);

/* The GSL error messages of the last edb200_get_loglike_matrix / edb200_emission call, as the reference prints them while it
 * computes (src/error.c:45-48: "ERROR <file> <line> <reason>" + "Default GSL error handler invoked." for every failing call,
 * then evaluation continues): per failing cell, in the reference's order (bin, then state), the error site(s) inside
 * gsl_sf_lnbeta (src/beta.c:44, 56, 59; src/VP_gamma.c:803, 1239, 1253, 1261, 1283) and the value wrapper's report
 * (src/beta.c:163).  Writes the text of events first_event, first_event + 1, ... as far as buflen allows (NUL-terminated),
 * sets *next_event to the first event not written, returns the number of failing cells the call had (the log keeps the
 * first 65,536).  csrc/r_glue.c prints it through Rprintf. */
EDB200_API int64_t edb200_gsl_error_log(char *buf, size_t buflen, int64_t first_event, int64_t *next_event);

/* Replaces C_hmm(nstates, nobs, transitions, probabilities, positions, expectedLength)  src/hmm.cpp:18-167.
 * transitions: double[S*S] column-major (R matrix; [k + S*j] = P(k -> j)); probabilities: double[nobs*S]
 * column-major in HMM state order (0 = normal); positions: int32[nobs].
 * path_out: int32[nobs]; calls_out: int32[4*call_cap] row-major (start.p, end.p, type, nexons), 1-based like
 * hmm.cpp:114-115; *ncalls_out = number of calls found.  nstates 2..7 are accepted (the reference accepts
 * only 3; the .Call glue keeps that check).  Host pointers. */
EDB200_API int edb200_hmm(int32_t nstates, int32_t nobs, const double *transitions, const double *probabilities,
               const int32_t *positions, double expected_length, int32_t *path_out, int32_t *calls_out,
               int32_t call_cap, int32_t *ncalls_out);

/* ---- cohort (batched) API: many samples over one shared bin set -------------------------------- */
/* Implements, for every sample, `new('ExomeDepth')`'s likelihood step + CallCNVs' per-chromosome Viterbi
 * with its framing (R/class_definition.R:342-374): dummy first/last observation, positions from bin starts,
 * likelihood columns permuted normal-first, start.p/end.p shifted by -1 and by the chromosome offset. */
typedef struct edb200_cohort edb200_cohort;

typedef struct edb200_cohort_spec {
    int64_t        n_bins;          /* bins, already ordered by (chromosome, midpoint) as CallCNVs orders them */
    int32_t        n_chains;        /* chromosomes */
    const int64_t *chain_offsets;   /* int64[n_chains+1], bins of chromosome c are [off[c], off[c+1]) */
    const int32_t *start;           /* int32[n_bins] bin start coordinates */
    const int32_t *end;             /* int32[n_bins] bin end coordinates   */
    int32_t        n_states;        /* 3 (reference), 5 or 7 (extensions)  */
    const double  *odds;            /* double[n_states] in likelihood-column (copy-number) order, or NULL = reference */
    double         mixture;         /* prop.tumor; used when odds == NULL  */
    const double  *transitions;     /* double[S*S] column-major in HMM order, or NULL = CallCNVs matrix from tp */
    double         transition_probability;   /* CallCNVs default 1e-4 */
    double         expected_cnv_length;      /* CallCNVs default 50000 */
    int32_t        skip_table_build;         /* 1: leave the log-transition table empty; the caller fills it with
                                                edb200_cohort_table_copy (ranks > 0 after an NCCL broadcast)     */
} edb200_cohort_spec;

/* Builds the framed positions and the log-transition table on the host (host libm, so that log/exp are the
 * reference's own bits) and uploads them. */
EDB200_API int  edb200_cohort_create(const edb200_cohort_spec *spec, edb200_cohort **out);
EDB200_API void edb200_cohort_destroy(edb200_cohort *c);

/* Per-cohort execution options.  Defaults (0 / -1 = auto) are what every documented number uses; the other values
 * exist so that tests can drive both sweep kernels over the same inputs and experiments can vary the placement.
 * Results never depend on them (tests/test_gpu_parity.py). */
#define EDB200_OPT_SWEEP       1   /* 0 auto; 1 one lane per (chain, state): any transition matrix (viterbi.cu);
                                      2 one thread per chain: CallCNVs-structured matrices, 3/5/7 states (viterbi_tpc.cu) */
#define EDB200_OPT_PARTS       2   /* chromosome groups a batch is pipelined over: 0 auto, 1 = one pass, .. 6 */
#define EDB200_OPT_VSPLIT      3   /* device-resident Viterbi as {longest chains} | {others} concurrently: -1 auto, 0, 1 */
#define EDB200_OPT_CRIT_WARPS  4   /* lane-per-state sweep: warps per CTA of the pass with the longest chains (1, 2, 4) */
#define EDB200_OPT_SWEEP_WARPS 5   /* sweep warps per CTA (lane-per-state: 4 or 8; thread-per-chain: 1..4), 0 auto */
#define EDB200_OPT_PACKPLAN    6   /* host pipeline: sweep packing per part as decimal digits, e.g. 122222; 0 auto */
#define EDB200_OPT_SEGMENTS    7   /* segmented Viterbi sweep (chains cut into concurrently swept, certified pieces): -1 auto, 0 off, 1 on */
#define EDB200_OPT_SEG_WARM    8   /* warm-up tiles (16 observations each) in front of a piece: 1 .. 64, 0 = default (2) */
#define EDB200_OPT_SEG_MIN     9   /* shortest piece in tiles, 0 = default (12); small values are for tests */
#define EDB200_OPT_SEG_REPAIR 10   /* test hook: 1 sends every chain through the repair pass of the segmented sweep */
#define EDB200_OPT_CHUNKS     12   /* host call with segmented sweeps: sample chunks the batch is pipelined over, 0 auto */
#define EDB200_OPT_RESERVE    11   /* host pipeline: SMs the emission launches of the later groups leave to the sweeps, 0 auto */
EDB200_API int  edb200_cohort_set_option(edb200_cohort *c, int option, int value);
/* What the last segmented sweep of this cohort did (synchronises the device): out[0] pieces the chains were cut into (0: the
 * last Viterbi pass was not segmented), out[1] decisions listed with a lead below 2^-14, out[2] chains sent to the repair
 * pass, out[3] (chain, sample) pairs among them, out[4..9] those pairs by reason: a NaN / Inf in a piece, the list of
 * decisions full, a seam with non-finite or huge values, a seam whose certified error outgrew its bound, a listed decision on
 * the final path, the test hook.  For tests and for bench.py's evidence; results never depend on it. */
EDB200_API int  edb200_cohort_segment_stats(edb200_cohort *c, int32_t out[10]);

/* Size in bytes and device address of the shared log-transition table (for an NCCL broadcast from rank 0). */
EDB200_API int  edb200_cohort_table(edb200_cohort *c, void **device_ptr, size_t *bytes);
/* Device-to-device copy of the table: direction 0 = table -> buf, 1 = buf -> table; enqueued on cuda_stream. */
EDB200_API int  edb200_cohort_table_copy(edb200_cohort *c, void *device_buf, int direction, void *cuda_stream);

typedef struct edb200_batch {
    int32_t        n_samples;
    const int32_t *observed;        /* int32[n_samples][obs_stride]  test counts                           */
    int64_t        obs_stride;      /* >= n_bins                                                           */
    const int32_t *reference;       /* int32[n_bins] shared aggregate (ref_stride 0) or [n_samples][ref_stride] */
    int64_t        ref_stride;
    const double  *phi;             /* double[n_samples]  over-dispersion per sample                       */
    const double  *expected;        /* double[n_samples]  expected proportion per sample                   */
    double        *ll;              /* out double[n_samples][n_states][ll_stride] (likelihood-column order); may be NULL in host mode */
    int64_t        ll_stride;       /* >= n_bins                                                           */
    int8_t        *path;            /* out int8[n_samples][path_stride], HMM states (0 normal), may be NULL */
    int64_t        path_stride;
    int32_t       *calls;           /* out int32[n_samples][call_cap][4]  (start.p, end.p, type, nexons)   */
    int32_t       *ncalls;          /* out int32[n_samples]                                                */
    int32_t        call_cap;
    /* CallCNVs post-processing (R/class_definition.R:393-400, :338), computed on the device so that the likelihood
     * matrix does not have to be copied to the host to be summed over the calls; both may be NULL */
    double        *call_stats;      /* out double[n_samples][call_cap][3] per call: sum over its bins of
                                       ll[,type] - ll[,normal] (BF before the log10(e) factor and signif),
                                       total*expected (reads.expected before as.integer), test (reads.observed) */
    double        *cor;             /* out double[n_samples]  cor(test, reference) over all bins              */
    /* Count-matrix ingestion layout (HOST mode only, used instead of `observed` when non-NULL): test
     * counts as uint16 [n_samples][obs16_stride] with 65535 standing for "see the overflow list", the list being
     * n_overflow entries sorted by flat index sample * n_bins + bin with their int32 counts.  Exome read counts per bin
     * fit 16 bits except for a handful of bins, so this halves the bytes that cross PCIe per call; the device widens
     * every chromosome group right behind its upload.  exomedepth_b200/cohort.py:pack_counts builds it from int32. */
    /* Per-bin fits (R/class_definition.R:121-147 `phi.bins > 1`, :168-180 covariate formulas): when per_bin_stride != 0,
     * phi and expected are double[n_samples][per_bin_stride] — one value per bin and sample, exactly the vectors
     * get_loglike_matrix receives — instead of one scalar per sample.  The per-state constants are then rebuilt per bin
     * (in-register kernel; the lattice kernels need per-sample constants). */
    int64_t        per_bin_stride;
    const uint16_t *observed16;
    int64_t        obs16_stride;
    int64_t        n_overflow;
    const int64_t *overflow_index;
    const int32_t *overflow_value;
    /* 12-bit ingestion layout (HOST mode only, used instead of `observed` / `observed16` when non-NULL): every row a
     * little-endian bit stream of 12 bits per bin — bins 2i and 2i+1 in bytes 3i .. 3i+2: v0 | v1 << 12 — with 4095 standing for
     * "see the overflow list" (the same list as above: counts of 4095 and beyond).  obs12_stride: bytes from one sample's row
     * to the next, a multiple of 4 and >= 3 * ceil(n_bins / 2); rows 4-byte aligned.  A quarter fewer bytes over PCIe than the
     * 16-bit layout — what a host with eight GPUs uploading at once is bound by.  exomedepth_b200/cohort.py:pack_counts12. */
    const uint8_t *observed12;
    int64_t        obs12_stride;
} edb200_batch;

/* mode: 0 = auto (full lattice from 106,496 bins; panel lattice from 4,096 bins when samples x states >= 2 x SMs;
 * in-register otherwise), 1 = force in-register evaluation, 2 = force the full shared-memory lattice,
 * 3 = force the panel-sized lattice (cohort API only) */
#define EDB200_EMISSION_AUTO   0
#define EDB200_EMISSION_DIRECT 1
#define EDB200_EMISSION_TABLE  2
#define EDB200_EMISSION_PANEL  3

/* All pointers in `b` are DEVICE pointers on the selected GPU; work is enqueued on `cuda_stream`
 * (a cudaStream_t, 0 = default stream) and NOT synchronised.  ll must be non-NULL (the Viterbi reads it).
 * what: bit 0 emission, bit 1 Viterbi, bit 2 CallCNVs post-processing (call_stats / cor; needs the calls of a
 * Viterbi pass over the same batch, in this call or an earlier one). */
EDB200_API int edb200_cohort_run_device(edb200_cohort *c, const edb200_batch *b, int what, int emission_mode, void *cuda_stream);

/* Same, HOST pointers: copies in, runs, copies the non-NULL outputs back, synchronises. */
EDB200_API int edb200_cohort_run_host(edb200_cohort *c, const edb200_batch *b, int emission_mode);

/* ---- CUDA-graph replay of a device-resident batch ------------------------------------------------------
 * Small panels (BASELINE config 4: 512 samples x 5,000 bins x 7 states) are bound by the ~10 kernel launches and the
 * stream fork / join of edb200_cohort_run_device, not by the kernels.  edb200_cohort_capture_device runs the batch
 * once on an internal stream (so that every workspace is sized), records the same enqueue into a CUDA graph and
 * returns it; edb200_graph_launch replays it on cuda_stream without synchronising.  The graph holds the ADDRESSES in
 * `b` and of the library's workspaces: the batch's device buffers must stay allocated, their contents may change
 * between replays (that is the point).  A replay fails with EDB200_ERR_ARG after the cohort was destroyed or after
 * any library workspace was re-allocated (a larger batch ran since): capture again.  Not available while
 * edb200_profile is on.  Replaces the per-sample R loop around the two .Call routines (R/class_definition.R:184-189,
 * R/tools.R:97) for callers that process many same-shaped batches. */
typedef struct edb200_graph edb200_graph;
EDB200_API int  edb200_cohort_capture_device(edb200_cohort *c, const edb200_batch *b, int what, int emission_mode,
                                             edb200_graph **out);
EDB200_API int  edb200_graph_launch(edb200_graph *g, void *cuda_stream);
EDB200_API void edb200_graph_destroy(edb200_graph *g);

/* ---- forward pass and transition-probability grid (EXTENSION: the reference has neither) ----
 * For every sample and every transition probability tp_grid[g] (CallCNVs matrix of R/class_definition.R:343-347
 * built for that tp; tp_grid == NULL: the cohort's own matrix, n_grid ignored): the sum over chromosomes of the
 * forward log-likelihood with the same framing as the Viterbi (dummy first/last observation, forced end in the
 * normal state).  loglik: double[n_samples][n_grid]; best: int32[n_samples] index of the first maximiser (the
 * per-sample MLE over the grid), may be NULL.  Definition: oracle/oracle.c:edo_forward_loglik; parity unpinned. */
/* device pointers (b->ll filled by edb200_cohort_run_device; loglik, best on the device); tp_grid is a HOST pointer */
EDB200_API int edb200_cohort_forward_device(edb200_cohort *c, const edb200_batch *b, const double *tp_grid, int32_t n_grid,
                                            double *loglik, int32_t *best, void *cuda_stream);
/* host pointers; runs over the likelihoods the most recent edb200_cohort_run_host left resident in HBM */
EDB200_API int edb200_cohort_forward_last(edb200_cohort *c, const double *tp_grid, int32_t n_grid, double *loglik, int32_t *best);

/* ---- select.reference.set: the correlation sweep ---------------------------------------
 * Replaces R/optimize_reference_set.R:100 — cor(x/(bin.length*sum(x)/10^6), test/(bin.length*sum(test)/10^6)) of
 * every candidate against the test sample — for all samples of a cohort at once: with leave-one-out cohorts the bins
 * selected at :81-88 are the same for every test sample, so the sweep is one Pearson matrix over those bins.
 * counts: int32[n_samples][stride]; bin_length: double[n_bins] or NULL (all 1, :66); selected: the 0-based indices of
 * the selected bins (the filter itself is a few quantiles, done by the caller: exomedepth_b200/refset.py).
 * cor_out: double[n_rows][n_samples], row r = correlations of sample row0 + r against every sample (diagonal 1).
 * The beta-binomial refits of the greedy loop that follows in R (:113-141, aod::betabin) are out of scope. */
EDB200_API int edb200_refset_correlations(const int32_t *counts, int64_t stride, int32_t n_samples,
                                          const double *bin_length, const int32_t *selected, int64_t n_selected,
                                          int32_t row0, int32_t n_rows, double *cor_out);
/* The two stages on DEVICE pointers, enqueued on cuda_stream without synchronising (multi-GPU: every rank standardises
 * its own samples, the rows are all-gathered, every rank then forms its block of the matrix).
 * z: double[n_samples][k_pad], k_pad = edb200_refset_kpad(n_selected). */
EDB200_API int64_t edb200_refset_kpad(int64_t n_selected);
EDB200_API int edb200_refset_standardize_device(const int32_t *counts, int64_t stride, int32_t n_samples,
                                                const double *bin_length, const int32_t *selected, int64_t n_selected,
                                                double *z, void *cuda_stream);
/* cor_out[m][n] = rows of za against rows of zb */
EDB200_API int edb200_refset_gram_device(const double *za, int32_t m, const double *zb, int32_t n, int64_t n_selected,
                                         double *cor_out, void *cuda_stream);
/* The sharded sweep WITHOUT an all-gather (one process per GPU of one NVLink / NVSwitch box): every rank keeps its
 * standardised rows in a block owned by this library; the other ranks map it with CUDA IPC and the Gram kernel reads
 * the B tiles straight from their owners' memory while it multiplies others — the all-gather is fused into the
 * contraction and no rank holds the whole matrix.  Protocol (exomedepth_b200/shard.py:refset_sweep, fused=True):
 *   1. edb200_refset_block_alloc            allocate the block (rows_per_rank rows, zero filled), get its IPC handle
 *   2. exchange the handles (any channel), edb200_refset_peers_open(all handles, world, my_rank)
 *   3. edb200_refset_standardize_device into the block; synchronise the device; barrier across the ranks
 *   4. edb200_refset_gram_peers_device      this rank's m rows against all n_total rows (row j lives on rank
 *                                           j / rows_per_rank); identical bits to the all-gather form
 *   5. barrier; edb200_refset_peers_close */
#define EDB200_IPC_HANDLE_BYTES 64
EDB200_API int edb200_refset_block_alloc(int32_t rows_per_rank, int64_t n_selected, void **z_dev, void *ipc_handle_out);
EDB200_API int edb200_refset_peers_open(const void *handles /* world x 64 bytes, rank order */, int32_t world, int32_t my_rank);
EDB200_API int edb200_refset_peers_close(void);
EDB200_API int edb200_refset_gram_peers_device(int32_t m, int32_t rows_per_rank, int32_t n_total, int64_t n_selected,
                                               double *cor_out, void *cuda_stream);

/* ---- beta-binomial fit of new('ExomeDepth') ---------------------------------------------
 * Stands in for aod::betabin(cbind(test, reference) ~ 1, random = ~ 1, link = 'logit') + aod::fitted
 * (R/class_definition.R:118-119, 168) for the default formula: per sample the maximum-likelihood expected proportion
 * mu (= x@expected, constant over the bins) and over-dispersion phi (= x@phi) of
 * test ~ BetaBinomial(test + reference, a = mu(1-phi)/phi, b = (1-mu)(1-phi)/phi).  aod is third-party and absent from the
 * reference tree: parity unpinned, the result is the maximiser of the likelihood (gradient < 1e-12 of the curvature).
 * observed: int32[n_samples][obs_stride]; reference: int32[n_bins] (ref_stride 0) or [n_samples][ref_stride].
 * loglik: the maximised log-likelihood without the binomial coefficients; info[s] >= 0: Newton iterations,
 * -1: degenerate sample (no reads / all reads in the test), -2: negative counts, or more than 4096 bins beyond the
 * kernel's histogram caps (6143 test reads, 22527 reference or total reads in one bin; up to 4096 such bins per sample
 * are handled exactly), -3: iteration cap reached, -4: no over-dispersion (phi -> 0).
 * Returns EDB200_WARN_NAN when any sample has info < 0 (its mu, phi are NaN except for -3 / -4). */
EDB200_API int edb200_betabin_fit(const int32_t *observed, int64_t obs_stride, const int32_t *reference, int64_t ref_stride,
                                  int32_t n_samples, int64_t n_bins, double *mu, double *phi, double *loglik, int32_t *info);
/* same on DEVICE pointers, enqueued on cuda_stream, not synchronised (always returns 0 unless the launch fails) */
EDB200_API int edb200_betabin_fit_device(const int32_t *observed, int64_t obs_stride, const int32_t *reference, int64_t ref_stride,
                                         int32_t n_samples, int64_t n_bins, double *mu, double *phi, double *loglik,
                                         int32_t *info, void *cuda_stream);

/* get.power.betabinom(size, my.phi, my.p, my.alt.p) for n problems at once (R/tools.R:128-166, theory = FALSE, limit = FALSE:
 * the expected log10 Bayes factor of the alternative proportion over one beta-binomial draw of `size` reads) — the inner step
 * of select.reference.set's greedy loop (R/optimize_reference_set.R:136-140), one CTA per problem.  Host pointers. */
EDB200_API int edb200_power_betabinom(const int32_t *size, const double *phi, const double *p, const double *alt_p, int32_t n,
                                      double *expected_bf);

/* sticky status word of device-side warnings since the last call with reset != 0 (EDB200_WARN_*) */
EDB200_API int edb200_status(int reset);

#ifdef __cplusplus
}
#endif
#endif /* EXOMEDEPTH_B200_H */
