// CPU check of the structured one-thread Viterbi step (exomedepth_b200/csrc/viterbi_step.h) against the plain scan of
// src/hmm.cpp:66-88, compiled and run by tests/test_host_tables.py.  Chains are driven through both for thousands of
// steps on adversarial inputs: values from small discrete sets (exact ties in V and in the candidates), huge emissions
// (distinct V that round to one candidate), -Inf and NaN emissions, -Inf transition terms, real CallCNVs rows built by
// host_tables.cpp from hg19-like gaps (incl. zero and negative ones).  V must agree bit for bit and the back-pointers exactly.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "host_tables.h"
#include "viterbi_step.h"

static int fails = 0;
static uint64_t bits(double x)
{
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

template <int S>
static void expand_row(double c0, double c1, const edb::StructRow& r, double* lt)
{
    for (int j = 0; j < S; j++)
        for (int k = 0; k < S; k++)
            lt[j * S + k] = k == 0 ? (j == 0 ? c0 : c1) : j == 0 ? r.b0 : k == j ? r.sf : r.ot;
}

static unsigned hi_word(double x) { return (unsigned)(bits(x) >> 32); }

template <int S>
static void run_case(int mode, uint64_t seed, int steps, long long* slow_hits, long long* fast_steps, long long* resolved, long long* spec_steps)
{
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const double ninf = -HUGE_VAL;
    double Va[S], Vb[S];
    for (int j = 0; j < S; j++) Va[j] = Vb[j] = j == 0 ? 0.0 : ninf;
    const double tp = mode == 3 ? 0.3 : 1e-4;
    const double c0 = std::log(1 - tp), c1 = std::log(tp / (S - 1));
    for (int i = 0; i < steps; i++) {
        edb::StructRow r{};
        double em[S];
        const double d = mode == 2 ? (U(rng) < 0.1 ? 1.0 : (U(rng) < 0.1 ? 1.0000426 : U(rng))) : std::exp(-U(rng) * 3);
        const double t_b0 = d * 0.5 + (1.0 - d) * (1 - tp), t_sf = d * 0.5 + (1.0 - d) * (tp / (S - 1)), t_ot = d * 0.0 + (1.0 - d) * (tp / (S - 1));
        r.b0 = std::log(t_b0);
        r.sf = std::log(t_sf);
        r.ot = std::log(t_ot);                      // -Inf at d = 1, NaN beyond (stored as -Inf like the device table)
        if (r.ot != r.ot) r.ot = ninf;
        if (r.b0 != r.b0) r.b0 = ninf;
        if (r.sf != r.sf) r.sf = ninf;
        for (int j = 0; j < S; j++) {
            switch (mode) {
                case 0: em[j] = -(double)(rng() % 3); break;                          // tiny discrete set: ties everywhere
                case 1: em[j] = -U(rng) * 20 - (j ? 0 : U(rng) < 0.05 ? 30 : 0); break;   // realistic, CNV states win now and then
                case 2: em[j] = U(rng) < 0.02 ? ninf : U(rng) < 0.02 ? NAN : -(double)(rng() % 5) * 0.25; break;
                case 3: em[j] = (U(rng) < 0.3 ? -1e15 : 0.0) - (double)(rng() % 4) * 1e-3; break;   // huge magnitudes: roundings collapse
                default: em[j] = j == 0 ? -5.0 - U(rng) : -U(rng); break;           // CNV states lead for long stretches
            }
        }
        if (mode == 0 && (i % 7) == 0) { r.b0 = r.sf = r.ot = -(double)(rng() % 2); }
        if (mode == 6) {                            // group candidate and k = 0 candidate within a few ulps of each other at 2^39
            const double base = -(3.0e8 + U(rng) * 2.0e8), u = 5.9604644775390625e-08;
            const int ks = 1 + (int)(rng() % (S - 1));
            for (int j = 0; j < S; j++) {
                Va[j] = Vb[j] = j == 0 ? base : base - 40.0 - (double)(rng() % 5);
                em[j] = U(rng) < 0.5 ? -(double)(rng() % 7) * u * 0.5 : -U(rng) * 1.0e8;
            }
            const double lt = (rng() & 1) ? r.ot : r.b0, cc = (lt == r.ot) ? c1 : c0;
            // even steps: at the candidates' tie; odd steps: at the edge of the speculative step's acceptance test (margin 2^-12) —
            // with a margin of 0 this placement makes the speculative step accept ~50 wrong steps per million
            if (lt > -1e300) Va[ks] = Vb[ks] = (base + (cc - lt) - (i & 1 ? edb::kSpecMargin : 0.0)) + ((double)(rng() % 9) - 4.0) * u;
        }
        if (mode == 5) {                            // the speculative step's margin: magnitudes just below 2^40, states within a few units
            const double base = -U(rng) * 1.0e9, eb = U(rng) < 0.5 ? -U(rng) * 1.0e9 : -U(rng) * 3;
            for (int j = 0; j < S; j++) {
                Va[j] = Vb[j] = base + (j == 0 ? 0.0 : -(double)(rng() % 64) * 0.25 + (U(rng) < 0.3 ? 8.0 : -6.0));
                em[j] = eb - (double)(rng() % 8) * 0.125;
            }
        }
        if (mode == 3) {                            // collapse: keep V small next to the emissions
            for (int j = 0; j < S; j++)
                if (Va[j] < -1e17) { Va[j] = Vb[j] = -(double)(rng() % 3) * 1e-3; }
        }
        double lt[S * S];
        expand_row<S>(c0, c1, r, lt);
        unsigned aa[S], ab[S];
        double Vprev[S];
        for (int j = 0; j < S; j++) Vprev[j] = Va[j];
        // the kernel's dispatch: the checked step when a special value is in sight, else the branch-free step plus
        // the deferred index scan of the destinations it flags
        bool special = edb::nonfinite_hi(hi_word(Va[0]));
        for (int j = 0; j < S; j++) special = special || edb::nonfinite_hi(hi_word(em[j]));
        // even seeds divisible by 4: the speculative step first, as the kernel does (single step here, pairs there)
        bool spec_ok = false;
        // (the speculative step relies on the row property build_struct_rows verifies; the synthetic rows of mode 0 break it)
        const bool row_ok = r.ot == ninf || (r.b0 > ninf && r.ot - c1 <= r.b0 - c0 + 9.5367431640625e-07);
        if (!special && row_ok && (seed & 3) == 0) {
            bool big = edb::big_or_nonfinite_hi(hi_word(Va[0]));
            for (int j = 0; j < S; j++) big = big || edb::big_or_nonfinite_hi(hi_word(em[j]));
            if (!big) {
                double Vs[S];
                unsigned as[S];
                for (int j = 0; j < S; j++) Vs[j] = Va[j];
                bool okk = true;
                const unsigned bits = edb::viterbi_step_spec<S>(Vs, em, c0, c1, c0 - edb::kSpecMargin, r, okk);
                as[0] = 0;
                for (int j = 1; j < S; j++) as[j] = (bits >> (j - 1) & 1u) ? (unsigned)j : 0u;
                if (okk) {
                    spec_ok = true;
                    (*spec_steps)++;
                    for (int j = 0; j < S; j++) { Va[j] = Vs[j]; aa[j] = as[j]; }
                }
            }
        }
        if (spec_ok) {}
        else if (special || (seed & 1)) edb::viterbi_step_struct<S>(Va, em, c0, c1, r, aa);
        else {
            const unsigned need = edb::viterbi_step_fast<S>(Va, em, c0, c1, r, aa);
            for (int j = 0; j < S; j++)
                if (need >> j & 1) { aa[j] = edb::viterbi_resolve_arg<S>(Vprev, em[j], j, c0, c1, r); (*resolved)++; }
            (*fast_steps)++;
        }
        edb::viterbi_step_scan<S>(Vb, em, lt, ab);
        for (int j = 0; j < S; j++) {
            if (bits(Va[j]) != bits(Vb[j]) || aa[j] != ab[j]) {
                if (fails < 10) {
                    printf("FAIL S=%d mode=%d step=%d j=%d  V %.17g vs %.17g  arg %u vs %u\n  Vprev:", S, mode, i, j, Va[j], Vb[j], aa[j], ab[j]);
                    for (int k = 0; k < S; k++) printf(" %.17g", Vprev[k]);
                    printf("\n  em:");
                    for (int k = 0; k < S; k++) printf(" %.17g", em[k]);
                    printf("\n  c0 %.17g c1 %.17g b0 %.17g sf %.17g ot %.17g\n", c0, c1, r.b0, r.sf, r.ot);
                }
                fails++;
                Va[j] = Vb[j];
            }
            if (ab[j] != 0 && ab[j] != (unsigned)j && ab[j] != 7u) (*slow_hits)++;
        }
    }
}

// real rows: the structured values must be what host_tables.cpp puts into the general table
template <int S>
static void check_real_rows()
{
    const int n = 4000;
    std::mt19937_64 rng(7);
    std::vector<int32_t> pos(n);
    int32_t p = 100000;
    for (int i = 0; i < n; i++) {
        const int g = (int)(rng() % 100);
        p += g < 3 ? 0 : g < 6 ? -(int32_t)(rng() % 2000) : g < 90 ? (int32_t)(rng() % 4000) : (int32_t)(rng() % 400000);
        pos[i] = p;
    }
    double T[S * S];
    edb::callcnvs_transitions(S, 1e-4, T);
    const int pitch = edb::lt_pitch(S), js = edb::lt_jstride(S);
    std::vector<double> lt((size_t)(n + 16) * pitch, 0.0);
    edb::build_log_transition_rows(S, T, pos.data(), n, 50000.0, lt.data(), pitch);
    edb::nan_to_neg_inf(lt.data(), lt.size());
    std::vector<edb::StructRow> rows(n);
    double c0 = 0, c1 = 0;
    const int ok = edb::build_struct_rows(S, lt.data(), pitch, n, rows.data(), &c0, &c1);
    if (!ok) { printf("FAIL: CallCNVs rows not recognised as structured (S=%d)\n", S); fails++; return; }
    for (int i = 1; i < n; i++)
        for (int j = 0; j < S; j++)
            for (int k = 0; k < S; k++) {
                const double want = lt[(size_t)i * pitch + j * js + k];
                const double got = k == 0 ? (j == 0 ? c0 : c1) : j == 0 ? rows[i].b0 : k == j ? rows[i].sf : rows[i].ot;
                if (bits(want) != bits(got)) { if (fails < 10) printf("FAIL row %d j %d k %d\n", i, j, k); fails++; }
            }
    // an arbitrary matrix is refused
    T[1 + S * 2] = 0.01;
    edb::build_log_transition_rows(S, T, pos.data(), n, 50000.0, lt.data(), pitch);
    edb::nan_to_neg_inf(lt.data(), lt.size());
    if (edb::build_struct_rows(S, lt.data(), pitch, n, rows.data(), &c0, &c1)) { printf("FAIL: arbitrary matrix accepted as structured\n"); fails++; }
}

int main()
{
    long long hits = 0, fast = 0, resolved = 0, spec = 0;
    for (int mode = 0; mode < 7; mode++)
        for (uint64_t seed = 1; seed <= 8; seed++) {       // odd seeds: the checked step throughout; even: the kernel's dispatch
            const int steps = mode == 6 && (seed & 3) == 0 ? 400000 : 20000;
            run_case<3>(mode, seed, steps, &hits, &fast, &resolved, &spec);
            run_case<5>(mode, seed * 11, steps, &hits, &fast, &resolved, &spec);
            run_case<7>(mode, seed * 13, steps, &hits, &fast, &resolved, &spec);
        }
    printf("branch-free steps: %lld, destinations resolved by the deferred scan: %lld\n", fast, resolved);
    printf("speculative steps accepted: %lld\n", spec);
    if (fast < 100000 || resolved < 10000 || spec < 20000) { printf("FAIL: the branch-free path is not exercised\n"); fails++; }
    check_real_rows<3>();
    check_real_rows<5>();
    check_real_rows<7>();
    printf("steps with a back-pointer into another CNV state: %lld\n", hits);
    if (hits < 1000) { printf("FAIL: the inputs do not exercise the group paths\n"); fails++; }
    printf(fails ? "FAILED %d\n" : "ok\n", fails);
    return fails ? 1 : 0;
}
