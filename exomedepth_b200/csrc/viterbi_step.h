// viterbi_step.h — one Viterbi step of a whole chain in ONE thread, for transition matrices with the CallCNVs
// structure (R/class_definition.R:343-347 and its S-state generalisation, host_tables.cpp:callcnvs_transitions).
//
// The reference's step (src/hmm.cpp:66-88) scans, per destination state j, the S candidates
//     cand_k = (em_j + V[k]) + log t(k -> j)          in that association, k = 0 .. S-1,
// keeps the FIRST maximum (strict '>'), leaves from = -1 when no candidate beats -Inf and forces from = 0
// when the emission is -Inf.  With the CallCNVs matrix log t(k -> j) takes at most three values per destination
// and observation (host_tables.h: StructRow):
//     j = 0:  k = 0 -> c0,  k > 0 -> b0
//     j > 0:  k = 0 -> c1,  k = j -> sf,  every other k -> ot
// so the candidates of a group that shares its transition term are  f(V[k])  for one monotone function
// f(x) = fl(fl(em_j + x) + lt): rounding is monotone, hence  max_k f(V[k]) = f(max_k V[k])  EXACTLY, and the group
// needs two additions instead of two per member.  What monotonicity does not give is the reference's winner INDEX
// when a group wins: the lowest k whose candidate equals the maximum, which can be a member with a smaller V
// whose candidate rounds to the same value.  Those cases — a group of "other" copy-number states winning or
// tying, or the two largest V of the return-to-normal group rounding to one candidate — take an exact scan of
// the group (rare, divergent; the value is already exact, only the index is settled there).
// Everything is IEEE add / compare: bit-identical to the reference's scan (tests/native/viterbi_step_check.cpp
// checks it against the plain scan on adversarial inputs; the GPU parity tests against the reference itself).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define EDB_STEP_HD __host__ __device__ __forceinline__
#else
#define EDB_STEP_HD inline
#endif

namespace edb {

#if defined(__CUDA_ARCH__)
#define EDB_ADD(a, b) __dadd_rn((a), (b))
#else
#define EDB_ADD(a, b) ((a) + (b))          // host build: -ffp-contract=off (additions only; nothing to contract anyway)
#endif

struct StructRow {      // the distance-dependent log-transition terms of one observation (host libm, host_tables.cpp)
    double b0;          // log t(k -> 0), k > 0
    double sf;          // log t(j -> j), j > 0
    double ot;          // log t(k -> j), j > 0, k not in {0, j}
    double pad;
};

// top two VALUES of V[1..S-1] and the lowest index holding the largest (the reference's first maximum)
template <int S>
struct Top2 {
    double m1, m2;
    int i1;
};

template <int S>
EDB_STEP_HD Top2<S> top2_cnv(const double* V)
{
    Top2<S> r;
    if (S == 3) {
        const bool a = V[2] > V[1];
        r.m1 = a ? V[2] : V[1];
        r.m2 = a ? V[1] : V[2];
        r.i1 = a ? 2 : 1;
    } else {
        // pairs (1,2), (3,4) [, (5,6)]: winner, loser, winner's index; then merges of (m1, i1, m2) records
        double w[(S - 1) / 2], l[(S - 1) / 2];
        int ix[(S - 1) / 2];
#pragma unroll
        for (int p = 0; p < (S - 1) / 2; p++) {
            const double x = V[2 * p + 1], y = V[2 * p + 2];
            const bool a = y > x;
            w[p] = a ? y : x;
            l[p] = a ? x : y;
            ix[p] = a ? 2 * p + 2 : 2 * p + 1;
        }
        double m1 = w[0], m2 = l[0];
        int i1 = ix[0];
#pragma unroll
        for (int p = 1; p < (S - 1) / 2; p++) {
            const bool c = w[p] > m1;
            const double A = w[p] > m2 ? w[p] : m2;          // second when the left record keeps the lead
            const double B = m1 > l[p] ? m1 : l[p];          // second when the right pair takes it
            m2 = c ? B : A;
            i1 = c ? ix[p] : i1;
            m1 = c ? w[p] : m1;
        }
        r.m1 = m1;
        r.m2 = m2;
        r.i1 = i1;
    }
    return r;
}

// One step for all S destination states.  V: in/out.  em: emissions in HMM state order (0 = normal), as stored
// (may be NaN or -Inf).  arg[j]: the reference's from_where (0..S-1), 7 for its -1.
template <int S>
EDB_STEP_HD void viterbi_step_struct(double* V, const double* em, double c0, double c1, const StructRow& row, unsigned* arg)
{
    static_assert(S == 3 || S == 5 || S == 7, "structured step: 3, 5 or 7 states");
    const double ninf = -HUGE_VAL;
    const Top2<S> t2 = top2_cnv<S>(V);
    double nv[S];
    // ---- destination 0: k = 0 with c0, every k > 0 with b0
    {
        const double e = em[0] != em[0] ? ninf : em[0];
        const double cand0 = EDB_ADD(EDB_ADD(e, V[0]), c0);
        const double g1 = EDB_ADD(EDB_ADD(e, t2.m1), row.b0);
        const double g2 = EDB_ADD(EDB_ADD(e, t2.m2), row.b0);
        const bool p0 = g1 > cand0;
        const double best = p0 ? g1 : cand0;
        unsigned a = p0 ? (unsigned)t2.i1 : 0u;
        if (p0 && g2 >= g1) {                       // rare: a lower-index state may round to the same candidate
            a = (unsigned)(S - 1);
#pragma unroll
            for (int k = S - 2; k >= 1; k--)
                if (EDB_ADD(EDB_ADD(e, V[k]), row.b0) == g1) a = (unsigned)k;
        }
        if (!(best > ninf)) a = 7u;
        if (em[0] == ninf) a = 0u;                  // hmm.cpp:87 (the stored emission: a NaN is not -Inf)
        nv[0] = best;
        arg[0] = a;
    }
    // ---- destinations j > 0: k = 0 with c1, k = j with sf, the others with ot
#pragma unroll
    for (int j = 1; j < S; j++) {
        const double e = em[j] != em[j] ? ninf : em[j];
        const double cand0 = EDB_ADD(EDB_ADD(e, V[0]), c1);
        const double self = EDB_ADD(EDB_ADD(e, V[j]), row.sf);
        const double mo = S == 3 ? V[3 - j] : (t2.i1 == j ? t2.m2 : t2.m1);     // largest V among the other CN states
        const double oth = EDB_ADD(EDB_ADD(e, mo), row.ot);
        const bool p1 = self > cand0;
        double best = p1 ? self : cand0;
        unsigned a = p1 ? (unsigned)j : 0u;
        if (oth >= best) {                          // uncommon: a direct change between two CNV states wins or ties
            // lowest k outside {0, j} whose candidate equals the group's maximum
            unsigned ko = 7u;
#pragma unroll
            for (int k = S - 1; k >= 1; k--)
                if (k != j && EDB_ADD(EDB_ADD(e, V[k]), row.ot) == oth) ko = (unsigned)k;
            if (oth > best) {
                best = oth;
                a = ko;
            } else if (a != 0u && ko < a) a = ko;   // tie with k = j: the lower index was scanned first
        }
        if (!(best > ninf)) a = 7u;
        if (em[j] == ninf) a = 0u;
        nv[j] = best;
        arg[j] = a;
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
}

// ---- branch-free form for the sweep kernel ------------------------------------------------------------------
// viterbi_step_fast is the step above without its rare paths, as ONE straight-line block (the kernel runs one warp
// per SM sub-partition: every branch ends a scheduling region and exposes the full latency of what precedes it —
// with a branch per destination the step took ~600 cycles, profiles/r2b_tpc_first.txt).  Preconditions, checked by
// the caller per pair of observations with integer tests on the high words: no emission of the step is NaN or
// +-Inf, V[0] > -Inf, c0 and c1 finite — then every destination has a finite candidate from k = 0, so no
// "from = -1" and no forced 0.  It returns the exact V and, per destination, the winner index when that index is
// settled by the values alone; bit j of the returned mask is set when destination j needs viterbi_resolve_arg
// (a group of states that share a transition term won or tied: the index is the lowest member whose candidate
// equals the group's, which takes a scan of the members against the V of BEFORE the step).
template <int S>
EDB_STEP_HD unsigned viterbi_step_fast(double* V, const double* em, double c0, double c1, const StructRow& row, unsigned* arg)
{
    static_assert(S == 3 || S == 5 || S == 7, "structured step: 3, 5 or 7 states");
    double om[S];                                   // om[j], j > 0: largest V among the CN states other than j
    double m1;                                      // largest V[k], k > 0, and the lowest index holding it
    int i1;
    if (S == 3) {
        const bool a = V[2] > V[1];
        m1 = a ? V[2] : V[1];
        i1 = a ? 2 : 1;
        om[1] = V[2];
        om[2] = V[1];
    } else {
        double pm[(S - 1) / 2];                     // pair maxima (1,2), (3,4) [, (5,6)]
        int pi[(S - 1) / 2];
#pragma unroll
        for (int p = 0; p < (S - 1) / 2; p++) {
            const bool a = V[2 * p + 2] > V[2 * p + 1];
            pm[p] = a ? V[2 * p + 2] : V[2 * p + 1];
            pi[p] = a ? 2 * p + 2 : 2 * p + 1;
        }
        if (S == 5) {
            om[1] = V[2] > pm[1] ? V[2] : pm[1];
            om[2] = V[1] > pm[1] ? V[1] : pm[1];
            om[3] = V[4] > pm[0] ? V[4] : pm[0];
            om[4] = V[3] > pm[0] ? V[3] : pm[0];
            const bool c = pm[1] > pm[0];
            m1 = c ? pm[1] : pm[0];
            i1 = c ? pi[1] : pi[0];
        } else {
            const double m01 = pm[1] > pm[0] ? pm[1] : pm[0], m02 = pm[2] > pm[0] ? pm[2] : pm[0], m12 = pm[2] > pm[1] ? pm[2] : pm[1];
            om[1] = V[2] > m12 ? V[2] : m12;
            om[2] = V[1] > m12 ? V[1] : m12;
            om[3] = V[4] > m02 ? V[4] : m02;
            om[4] = V[3] > m02 ? V[3] : m02;
            om[5] = V[6] > m01 ? V[6] : m01;
            om[6] = V[5] > m01 ? V[5] : m01;
            const bool c = pm[1] > pm[0];
            const double t = c ? pm[1] : pm[0];
            const int ti = c ? pi[1] : pi[0];
            const bool c2 = pm[2] > t;
            m1 = c2 ? pm[2] : t;
            i1 = c2 ? pi[2] : ti;
        }
    }
    double nv[S];
    unsigned need = 0;
    {
        const double cand0 = EDB_ADD(EDB_ADD(em[0], V[0]), c0);
        const double g1 = EDB_ADD(EDB_ADD(em[0], m1), row.b0);
        const bool p0 = g1 > cand0;
        nv[0] = p0 ? g1 : cand0;
        arg[0] = p0 ? (unsigned)i1 : 0u;
        need |= p0 ? 1u : 0u;
    }
#pragma unroll
    for (int j = 1; j < S; j++) {
        const double cand0 = EDB_ADD(EDB_ADD(em[j], V[0]), c1);
        const double self = EDB_ADD(EDB_ADD(em[j], V[j]), row.sf);
        const double oth = EDB_ADD(EDB_ADD(em[j], om[j]), row.ot);
        const bool p1 = self > cand0;
        const double t = p1 ? self : cand0;
        const bool q = oth >= t;
        nv[j] = q ? oth : t;                        // equal values either way when they tie
        arg[j] = p1 ? (unsigned)j : 0u;
        need |= q ? (1u << j) : 0u;
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
    return need;
}

// Winner index of destination j by the reference's own scan (first maximum), from the V of before the step.
// Same preconditions as viterbi_step_fast (a finite candidate exists), so the result is never "from = -1".
template <int S>
EDB_STEP_HD unsigned viterbi_resolve_arg(const double* Vprev, double em_j, int j, double c0, double c1, const StructRow& row)
{
    double best = EDB_ADD(EDB_ADD(em_j, Vprev[0]), j == 0 ? c0 : c1);
    unsigned a = 0u;
#pragma unroll
    for (int k = 1; k < S; k++) {
        const double c = EDB_ADD(EDB_ADD(em_j, Vprev[k]), j == 0 ? row.b0 : (k == j ? row.sf : row.ot));
        if (c > best) {
            best = c;
            a = (unsigned)k;
        }
    }
    return a;
}

// ---- speculative form: the common case in ~30 FP64 instructions on a ~35-cycle dependent chain -------------------
// Away from CNV regions every destination is won by k = 0 or by k = j: then V'[0] = cand0 and, for j > 0,
// V'[j] = max(cand0_j, self_j) — two additions and one compare deep — and the group candidates only have to be
// shown to LOSE, which one comparison per step does for all destinations at once, conservatively:
//     (M1 + b0) < (V[0] + c0) - mu,      M1 = max_{k>0} V[k],  mu = 2^-12
// (the same inequality with (ot, c1) in place of (b0, c0) follows from it: the table builder verifies
// ot - c1 <= b0 - c0 + 2^-20 for every row, host_tables.cpp:build_struct_rows).
// Why that suffices.  Preconditions (the caller checks the high words): |V[0]|, |em| < 2^30; c0, c1 finite, lt <= 0.
// If M1 + lt < -2^33 the group candidate is below -2^32 and cand0 above it: nothing to show.  Otherwise every
// quantity involved is below 2^35 in magnitude, so each rounding moves a value by at most 2^-19: the test,
// evaluated in floating point, implies (M1 + lt) - (V[0] + c) < -mu + 2^-17 exactly, and the reference's
// floating-point candidates  fl(fl(em + V[k]) + lt),  fl(fl(em + V[0]) + c)  differ from their exact values by at
// most 2^-18 each; with V[k] <= M1 (and the 2^-20 of the row check) the group candidate is STRICTLY below cand0 <= the step's maximum:
// neither V' nor the winner index can involve it.  Returns false when a comparison fails (the pair of observations
// is then redone by the exact step; V and arg are garbage).  c0m = c0 - mu.
constexpr double kSpecMargin = 0.000244140625;      // 2^-12
// Returns the winner indices as ONE BIT per destination j > 0 (bit j-1 set: k = j won, clear: k = 0; destination 0 is
// always won by k = 0 here), and clears `ok` when the comparison fails.  The comparison is taken state by state
// (V[k] + b0 < V[0] + c0 - mu for every k > 0, which is the same statement as for their maximum): additions and
// compares only, no selects on the dependent chain.
template <int S>
EDB_STEP_HD unsigned viterbi_step_spec(double* V, const double* em, double c0, double c1, double c0m, const StructRow& row, bool& ok)
{
    static_assert(S == 3 || S == 5 || S == 7, "structured step: 3, 5 or 7 states");
    const double y = EDB_ADD(V[0], c0m);
    bool fine = ok;
#pragma unroll
    for (int k = 1; k < S; k++) fine = fine && (EDB_ADD(V[k], row.b0) < y);
    ok = fine;
    double nv[S];
    unsigned bits = 0u;
    nv[0] = EDB_ADD(EDB_ADD(em[0], V[0]), c0);
#pragma unroll
    for (int j = 1; j < S; j++) {
        const double cand0 = EDB_ADD(EDB_ADD(em[j], V[0]), c1);
        const double self = EDB_ADD(EDB_ADD(em[j], V[j]), row.sf);
        const bool p1 = self > cand0;
        nv[j] = p1 ? self : cand0;
        bits |= p1 ? (1u << (j - 1)) : 0u;
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
    return bits;
}

// ---- segments: steps that also certify their decisions (viterbi_seam.h) -------------------------------------------
// A chain cut into segments is swept from an approximate start vector: the values X of such a sweep follow the
// reference's values R up to a constant and a small error, |X[k] + C - R[k]| <= eps for every state k (viterbi_seam.h
// derives eps).  Every candidate then obeys |xc_k + C - rc_k| <= eps' (eps plus the roundings of the step), so a decision
// whose winner leads every other candidate by more than 2 eps' in X arithmetic has the same winner in the reference's.
// The steps below report the decisions whose lead is below kSegTau; the caller guarantees 2 eps' < kSegTau.
constexpr double kSegTau = 6.103515625e-05;             // 2^-14
constexpr unsigned kSegTauHi = (1023u - 14u) << 20;     // its high word

EDB_STEP_HD unsigned f64_hi(double x)
{
#if defined(__CUDA_ARCH__)
    return (unsigned)__double2hiint(x);
#else
    unsigned long long u;
    __builtin_memcpy(&u, &x, 8);
    return (unsigned)(u >> 32);
#endif
}

// viterbi_step_spec with the lead of every decision it takes: min_hi collects the smallest high word of |self - cand0|
// over the destinations (the decision between k = j and k = 0; the group candidates are kept away by the 2^-12 of the
// acceptance test itself).  The caller rejects the pair when min_hi < kSegTauHi, i.e. some |self - cand0| < 2^-14.
template <int S>
EDB_STEP_HD unsigned viterbi_step_spec_m(double* V, const double* em, double c0, double c1, double c0m, const StructRow& row, bool& ok,
                                         unsigned& min_hi)
{
    static_assert(S == 3 || S == 5 || S == 7, "structured step: 3, 5 or 7 states");
    const double y = EDB_ADD(V[0], c0m);
    bool fine = ok;
#pragma unroll
    for (int k = 1; k < S; k++) fine = fine && (EDB_ADD(V[k], row.b0) < y);
    ok = fine;
    double nv[S];
    unsigned bits = 0u, mh = min_hi;
    nv[0] = EDB_ADD(EDB_ADD(em[0], V[0]), c0);
#pragma unroll
    for (int j = 1; j < S; j++) {
        const double cand0 = EDB_ADD(EDB_ADD(em[j], V[0]), c1);
        const double self = EDB_ADD(EDB_ADD(em[j], V[j]), row.sf);
        const bool p1 = self > cand0;
        nv[j] = p1 ? self : cand0;
        bits |= p1 ? (1u << (j - 1)) : 0u;
        const unsigned h = f64_hi(EDB_ADD(self, -cand0)) & 0x7FFFFFFFu;      // -Inf (self = -Inf) and NaN read as "large"
        mh = h < mh ? h : mh;
    }
    min_hi = mh;
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
    return bits;
}

// The plain scan of src/hmm.cpp:66-88 over the structured row, keeping the runner-up: arg[j] is the reference's first
// maximum (in the arithmetic of the V given), and bit j of the result is set when destination j's winner leads the best
// other candidate by less than kSegTau (ties included); lead[j], if asked for, is that lead.  Preconditions (caller): every V and emission finite, c0 and c1
// finite — so a finite candidate from k = 0 exists, no "from = -1", no forced 0.
template <int S>
EDB_STEP_HD unsigned viterbi_step_margin(double* V, const double* em, double c0, double c1, const StructRow& row, unsigned* arg,
                                         double* lead = nullptr)
{
    const double ninf = -HUGE_VAL;
    double nv[S];
    unsigned close = 0u;
#pragma unroll
    for (int j = 0; j < S; j++) {
        double best = ninf, second = ninf;
        unsigned a = 7u;
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double lt = k == 0 ? (j == 0 ? c0 : c1) : j == 0 ? row.b0 : k == j ? row.sf : row.ot;
            const double c = EDB_ADD(EDB_ADD(em[j], V[k]), lt);
            const bool win = c > best;
            const double loser = win ? best : c;            // a candidate equal to the leader becomes the runner-up: lead 0
            second = loser > second ? loser : second;
            best = win ? c : best;
            a = win ? (unsigned)k : a;
        }
        const double ld = EDB_ADD(best, -second);
        if (!(ld >= kSegTau)) close |= 1u << j;
        if (lead) lead[j] = ld;                             // (NaN when every candidate is -Inf: listed, lead unknown)
        nv[j] = best;
        arg[j] = a;
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
    return close;
}

// How far a segment's RELATIVE vector can be from the reference's (viterbi_seam.h): D = the spread max_k err[k] - min_k
// err[k] of the deviations err[k] = (R[k] - R[0]) - (X[k] - X[0]), as two multipliers, D <= a * e0 + b * 2 rho  (e0: the
// bound at the seam, rho: the rounding of one candidate in both arithmetics).  A step gives
// err'[j] - err'[j'] = err[w(j)] - err[w(j')] + (rho_j - rho_j')  with w(j) the winner of destination j:
//   kind 0  every destination has the same winner (certified): the old deviations cancel,   (a, b) -> (0, 1)
//   kind 1  anything else — also a listed decision, whose winner may differ in the reference's arithmetic (max is
//           1-Lipschitz: every R'[j] - X'[j] stays inside the old interval widened by rho):   (a, b) -> (a, b + 1)
// (saturating: a chain whose multipliers run away is refused by the check kernel).
EDB_STEP_HD void seg_err_step(unsigned& a, unsigned& b, int kind)
{
    a = kind == 0 ? 0u : a;
    b = kind == 0 ? 1u : (b < (1u << 24) ? b + 1u : b);
}
// kind of a step from its winners (arg[j], 7 = none) and the mask of listed decisions
template <int S>
EDB_STEP_HD int seg_err_kind(const unsigned* arg, unsigned close)
{
    bool same = close == 0u;
#pragma unroll
    for (int j = 1; j < S; j++) same = same && arg[j] == arg[0];
    return same ? 0 : 1;
}

// exponent field of x at least 1023 + 30 (|x| >= 2^30, Inf or NaN), from the high word alone
constexpr unsigned kSpecBigHi2 = (1023u + 30u) << 21;          // compared with hi << 1
EDB_STEP_HD bool big_or_nonfinite_hi(unsigned hi) { return (hi << 1) >= kSpecBigHi2; }

// true when x is NaN or +-Inf (exponent field all ones), from the high word alone
EDB_STEP_HD bool nonfinite_hi(unsigned hi) { return (hi << 1) >= 0xFFE00000u; }

// The plain scan of src/hmm.cpp:66-88 for an arbitrary row of S x S log-transition terms lt[j * S + k]
// (NaN terms already stored as -Inf).  Used by the checks and as the definition the structured step must match.
template <int S>
EDB_STEP_HD void viterbi_step_scan(double* V, const double* em, const double* lt, unsigned* arg)
{
    const double ninf = -HUGE_VAL;
    double nv[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        const double e = em[j] != em[j] ? ninf : em[j];
        double best = ninf;
        unsigned a = 7u;
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double c = EDB_ADD(EDB_ADD(e, V[k]), lt[j * S + k]);
            if (c > best) {
                best = c;
                a = (unsigned)k;
            }
        }
        if (em[j] == ninf) a = 0u;
        nv[j] = best;
        arg[j] = a;
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
}

}  // namespace edb
