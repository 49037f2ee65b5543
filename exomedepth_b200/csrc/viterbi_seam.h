// viterbi_seam.h — certifying a Viterbi chain that was swept as independent SEGMENTS (host + device).
//
// The sweep of src/hmm.cpp:58-90 is one dependent chain per (sample, chromosome): ~20,000 steps for chromosome 1, which
// bounds a 256-sample batch however many SMs there are.  (max, +) recurrences forget their start — once every survivor
// path passes through one node the scores before it only add a constant — so a chain can be cut into segments that are
// swept concurrently, each from a short warm-up that starts `warm` tiles before its first observation with V = (0, -Inf,
// ...).  The reference's RESULT, however, is defined by its own FP64 roundings, which depend on the absolute values; the
// segments are therefore swept speculatively and every decision they take is CERTIFIED afterwards, or the chain is swept
// again sequentially (the "repair" pass: the plain exact kernel).  Nothing here is probabilistic.
//
// Notation: R_i[k] the reference's values, X_i[k] a segment's values, C_i = R_i[0] - X_i[0],
// err_i[k] = (R_i[k] - R_i[0]) - (X_i[k] - X_i[0]) the deviation of the segment's RELATIVE vector and D_i = max_k err_i[k] -
// min_k err_i[k] its spread; only relative values decide a step.  A candidate is fl(fl(em + V[k]) + lt): two roundings of
// relative size 2^-53 on intermediate results of magnitude <= M, so in both arithmetics together
// rc_k - xc_k = C_i + err_i[k] + rho_k, |rho_k| <= rho = 2^-52 (M_X + M_R).  Hence:
//  (1) for two candidates of one destination, (rc_k - rc_k') - (xc_k - xc_k') lies within D_i + 2 rho: a decision whose
//      winner leads the other candidates by more than D_i + 2 rho in X arithmetic has the same (unique) winner in R
//      arithmetic.  The sweep lists every decision with a lead below kSegTau (viterbi_step.h); seam_advance() refuses a
//      segment unless D_i + 2 rho <= kSegTau / 4 at every step.
//  (2) with the winners w(j) equal in both arithmetics,  err_{i+1}[j] = err_i[w(j)] + rho_j - (err_i[w(0)] + rho_0).  When
//      every destination has the same winner — the normal case away from CNV regions, where every state is reached from
//      state 0, and inside a called region, where every state is reached from the called one — the old deviations
//      CANCEL: D_{i+1} <= 2 rho, however long the segment.  (The absolute values keep drifting apart, at one ulp of R per
//      step; they never enter a decision.)  In every other case, including a listed decision whose winner may differ
//      (max is 1-Lipschitz), D grows by at most 2 rho.  The sweep keeps the two multipliers of viterbi_step.h:
//      seg_err_step per lane; the check needs their maxima and their values at the segment's end.
// A listed decision matters only when the traceback reads it, i.e. when the final path — built from certified decisions
// alone as long as it meets no listed one — is in that destination state at that observation (viterbi_seg_check_kernel);
// then, and when a seam cannot be certified, the whole chain goes to the repair pass.
//
// The seam itself: segment s - 1, certified up to its end with spread D', hands over X'_end[k]; segment s arrives from
// its warm-up with X_in[k].  With m = max_k |(X_in[k] - X_in[0]) - (X'_end[k] - X'_end[0])| segment s starts with
// D_0 <= D' + 2 m.  m is at rounding level when the warm-up has coalesced (every state's value derives from the newest
// few observations), and the seam fails (repair) when it has not.
#pragma once
#include <cmath>

#include "viterbi_step.h"

namespace edb {

constexpr double kSegEpsMax = kSegTau / 4;          // largest certified spread: D + 2 rho <= tau / 4, well below a listed lead
constexpr double kSegMagMax = 536870912.0;          // 2^29: magnitudes the acceptance test of the speculative step was proved for

// why a chain goes to the repair pass (bits of its flag words; edb200_cohort_segment_stats counts them)
constexpr int kBadNonFinite = 1, kBadListFull = 2, kBadSeamValues = 4, kBadSeamError = 8, kBadOnPath = 16, kBadForced = 32;

struct SeamState {
    double eps;        // bound of the spread D of the relative deviations at the end of the segments verified so far
    double cabs;       // bound of |C| = |R[0] - X[0]|
    int bad;
};

// what a piece reports per lane besides its seam vectors (ViterbiArgs::seam_mag, kSeamWords words per lane)
constexpr int kSeamWords = 6;       // mag_v, mag_e, max_a, max_b, end_a, end_b
struct PieceErr {
    unsigned mag_v, mag_e;          // largest (high word << 1) of the piece's V (at every second observation) / emissions
    unsigned max_a, max_b;          // largest error multipliers any recorded decision of the piece was taken under
    unsigned end_a, end_b;          // the multipliers after its last observation
};

// magnitude bound from the largest (high word << 1) seen: |x| < 2^(exponent - 1022); +Inf for NaN / Inf
EDB_STEP_HD double mag_bound(unsigned hi2)
{
    const int e = (int)(hi2 >> 21);
    if (e >= 0x7FF) return HUGE_VAL;
    return ldexp(1.0, e - 1022);
}

// One seam.  x_in: the segment's V after its warm-up; x_prev: the previous segment's V after its last observation;
// pe: what the segment reported; n_steps: its observations.
template <int S>
EDB_STEP_HD void seam_advance(SeamState& s, const double* x_in, const double* x_prev, const PieceErr& pe, int n_steps)
{
    double m = 0.0, big = 0.0;
    bool finite = true;
#pragma unroll
    for (int k = 0; k < S; k++) {
        finite = finite && (f64_hi(x_in[k]) << 1) < 0xFFE00000u && (f64_hi(x_prev[k]) << 1) < 0xFFE00000u;
        const double d = fabs((x_in[k] - x_in[0]) - (x_prev[k] - x_prev[0]));
        m = d > m ? d : m;
        big = fabs(x_in[k]) > big ? fabs(x_in[k]) : big;
        big = fabs(x_prev[k]) > big ? fabs(x_prev[k]) : big;
    }
    const double bv = mag_bound(pe.mag_v), be = mag_bound(pe.mag_e);
    if (!finite || !(bv < kSegMagMax) || !(be < kSegMagMax)) {
        s.bad = kBadSeamValues;
        return;
    }
    const double kUlp = 2.220446049250313e-16;                              // 2^-52
    const double e0 = (s.eps + 2.0 * m + 16.0 * kUlp * big) * 1.0001;      // (the three subtractions behind m, rounded)
    // |C| at the seam, and its drift inside the segment: |C_{i+1} - C_i| <= D_i + rho, below kSegTau per step
    const double cabs = s.cabs + fabs(x_prev[0] - x_in[0]) * (1.0 + 4.0 * kUlp) + e0 + (double)n_steps * kSegTau;
    // values between two recorded observations exceed the recorded bound by at most one emission and one transition term
    // (|log t| <= 1024 for every finite term: capi.cu, ensure_struct)
    const double m_x = bv + 2.0 * be + 2048.0;
    const double m_r = m_x + cabs + 1.0;
    const double rho = kUlp * (m_x + m_r) * 1.0001;
    const double worst = (double)pe.max_a * e0 + ((double)pe.max_b + 2.0) * 2.0 * rho;
    s.eps = (double)pe.end_a * e0 + ((double)pe.end_b + 1.0) * 2.0 * rho;
    s.cabs = cabs;
    if (!(worst <= kSegEpsMax) || !(bv + cabs < kSegMagMax)) s.bad = kBadSeamError;
}

}  // namespace edb
