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
//      arithmetic.  The sweep lists every decision with a lead below kSegTau (viterbi_step.h) together with its lead;
//      seam_check() refuses a segment unless D_i + 2 rho <= kSegTau / 4 at every step, and returns the bound the segment
//      actually reached: a listed decision whose lead exceeds it is certified after all.
//  (2) with the winners w(j) equal in both arithmetics,  err_{i+1}[j] = err_i[w(j)] + rho_j - (err_i[w(0)] + rho_0).  When
//      every destination has the same winner — the normal case away from CNV regions, where every state is reached from
//      state 0, and inside a called region, where every state is reached from the called one — the old deviations
//      CANCEL: D_{i+1} <= 2 rho, however long the segment.  (The absolute values keep drifting apart, at one ulp of R per
//      step; they never enter a decision.)  In every other case, including a listed decision whose winner may differ
//      (max is 1-Lipschitz), D grows by at most 2 rho.  The sweep keeps the two multipliers of viterbi_step.h:
//      seg_err_step per lane; the check needs their maxima and their values at the segment's end.
// An uncertified decision matters only when the traceback reads it, i.e. when the final path — built from certified
// decisions alone as long as it meets no uncertified one — is in that destination state at that observation
// (viterbi_seg_check_kernel);
// then, and when a seam cannot be certified, the whole chain goes to the repair pass.
//
// The seam itself: segment s - 1, certified up to its end with spread D', hands over X'_end[k]; segment s arrives from
// its warm-up with X_in[k].  With m = max_k |(X_in[k] - X_in[0]) - (X'_end[k] - X'_end[0])| segment s starts with
// D_0 <= D' + 2 m, D' = (end_b + 1) 2 rho of the previous segment (a segment none of whose steps cancelled the deviation
// it started with is refused).  m is at rounding level when the warm-up has coalesced (every state's value derives from the newest
// few observations), and the seam fails (repair) when it has not.
#pragma once
#include <cmath>

#include "viterbi_step.h"

namespace edb {

constexpr double kSegEpsMax = kSegTau / 4;          // largest certified spread: D + 2 rho <= tau / 4, well below a listed lead
constexpr double kSegMagMax = 536870912.0;          // 2^29: magnitudes the acceptance test of the speculative step was proved for

// why a chain goes to the repair pass (bits of its flag words; edb200_cohort_segment_stats counts them)
constexpr int kBadNonFinite = 1, kBadListFull = 2, kBadSeamValues = 4, kBadSeamError = 8, kBadOnPath = 16, kBadForced = 32;

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

// Every seam is judged on its own (the check kernel runs one thread per seam and sample), which takes a bound of
// |C| = |R[0] - X[0]| that does not depend on the seams before it.  C changes at a seam by X'_end[0] - X_in[0] (plus the
// seam's deviation, below kSegTau / 4) and drifts inside a piece by at most D + rho < kSegTau per step; with bv_p the
// magnitude bound of piece p's values, |X'_end[0]| + |X_in[0]| <= bv_{p-1} + bv_p, so along the whole line
//     |C| <= sum over its pieces of  2 bv_p + (n_p + 1) kSegTau
// — each sweep warp adds its piece's share to the line's sum when the piece ends (piece_cabs_share).
EDB_STEP_HD double piece_cabs_share(unsigned mag_v, int n_steps)
{
    return 2.0 * mag_bound(mag_v) * (1.0 + 1e-9) + ((double)n_steps + 1.0) * kSegTau;
}

// rho of a piece: the rounding of one candidate in both arithmetics, 2^-52 (M_X + M_R).  Values between two recorded
// observations exceed the recorded bound by at most one emission and one transition term (|log t| <= 1024 for every finite
// term: capi.cu, ensure_struct); M_R = M_X + |C|.
EDB_STEP_HD double piece_rho(const PieceErr& pe, double cabs)
{
    const double kUlp = 2.220446049250313e-16;                              // 2^-52
    const double m_x = mag_bound(pe.mag_v) + 2.0 * mag_bound(pe.mag_e) + 2048.0;
    return kUlp * (2.0 * m_x + cabs + 1.0) * 1.0001;
}

// One seam.  x_in: the piece's V after its warm-up; x_prev: the previous piece's V after its last observation; pe / prev:
// what the two pieces reported (prev = null: the previous piece starts the chain, its values are the reference's own);
// cabs: the line's bound of |C|.  Returns 0 when every decision the piece did not list is certified, else the reason;
// *certified: the lead above which a decision of this piece is certified.
template <int S>
EDB_STEP_HD int seam_check(const double* x_in, const double* x_prev, const PieceErr& pe, const PieceErr* prev, double cabs,
                           double* certified = nullptr)
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
    if (!finite || !(bv < kSegMagMax) || !(be < kSegMagMax) || !(cabs < kSegMagMax)) return kBadSeamValues;
    // spread of the previous piece's deviations at its end: (end_b + 1) 2 rho once a step has cancelled what it started with
    double eps_prev = 0.0;
    if (prev) {
        if (prev->end_a != 0u || !(mag_bound(prev->mag_v) < kSegMagMax) || !(mag_bound(prev->mag_e) < kSegMagMax)) return kBadSeamError;
        eps_prev = ((double)prev->end_b + 1.0) * 2.0 * piece_rho(*prev, cabs);
    }
    const double kUlp = 2.220446049250313e-16;
    const double e0 = (eps_prev + 2.0 * m + 16.0 * kUlp * big) * 1.0001;   // (the three subtractions behind m, rounded)
    const double rho = piece_rho(pe, cabs);
    const double worst = (double)pe.max_a * e0 + ((double)pe.max_b + 2.0) * 2.0 * rho;
    if (!(worst <= kSegEpsMax) || !(bv + cabs < kSegMagMax)) return kBadSeamError;
    // every decision of the piece whose lead exceeds `worst` (>= D_i + 2 rho at every step) is certified — also a LISTED one:
    // the list holds the leads below kSegTau, most of which are far above the bound the piece actually reached
    if (certified) *certified = worst;
    return 0;
}

}  // namespace edb
