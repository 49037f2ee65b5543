// CPU check of the host-side shared metadata (exomedepth_b200/csrc/host_tables.cpp), compiled and run by
// tests/test_host_tables.py: the sweep placement for every CTA shape, the CallCNVs transition matrix
// (R/class_definition.R:343-347), the position framing (:368) and the log-transition rows against a direct
// evaluation of src/hmm.cpp:62-79.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <utility>
#include <vector>

#include "host_tables.h"

static int fails = 0;
#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } \
    } while (0)

static void check_schedule(const std::vector<int32_t>& nobs, int groups, int n_ctas, int W)
{
    std::vector<int32_t> begin, items;
    edb::viterbi_schedule(nobs.data(), (int)nobs.size(), groups, n_ctas, W, begin, items);
    const int n_slots = n_ctas * W;
    CHECK((int)begin.size() == n_slots + 1);
    CHECK(begin[0] == 0 && begin[n_slots] == (int)nobs.size() * groups);
    CHECK(items.size() == (size_t)2 * nobs.size() * groups);
    std::set<std::pair<int, int>> seen;
    int64_t busiest = 0, total = 0, longest = 0;
    for (int s = 0; s < n_slots; s++) {
        CHECK(begin[s] <= begin[s + 1]);
        int64_t load = 0;
        for (int q = begin[s]; q < begin[s + 1]; q++) {
            const int c = items[2 * q], g = items[2 * q + 1];
            CHECK(c >= 0 && c < (int)nobs.size() && g >= 0 && g < groups);
            CHECK(seen.insert({c, g}).second);             // every work item exactly once
            load += nobs[c];
        }
        if (load > busiest) busiest = load;
        total += load;
    }
    for (int32_t n : nobs) if (n > longest) longest = n;
    CHECK(seen.size() == nobs.size() * (size_t)groups);
    // greedy placement on the least loaded sub-partition, then on its less loaded warp: no warp carries more than the
    // mean load plus two work items
    CHECK((double)busiest <= (double)total / n_slots + 2.0 * (double)longest);
    // the longest work items sit alone on their sub-partition whenever there are at least as many sub-partitions as items
    if ((int64_t)nobs.size() * groups <= (int64_t)n_ctas * (W < 4 ? W : 4)) CHECK(busiest == longest);
}

// Shared-memory wavefronts of one warp-wide load under the bank model ncu confirmed on the S = 7 sweep
// (profiles/r1l_sweep_s7_summary.txt: 16 per LDS.128 with rows 64 bytes apart): 32 banks of 4 bytes; a load of `width`
// bytes per lane is served in phases of 128 / width lanes; within a phase, distinct words on one bank serialise and
// equal addresses are a broadcast.
static int wavefronts(const int* addr, int width)
{
    const int per_phase = 128 / width, words = width / 4;
    int total = 0;
    for (int l0 = 0; l0 < 32; l0 += per_phase) {
        int worst = 0;
        for (int bank = 0; bank < 32; bank++) {
            std::set<int> distinct;
            for (int l = l0; l < l0 + per_phase; l++)
                for (int w = 0; w < words; w++)
                    if ((addr[l] / 4 + w) % 32 == bank) distinct.insert(addr[l] / 4 + w);
            if ((int)distinct.size() > worst) worst = (int)distinct.size();
        }
        total += worst;
    }
    return total;
}

// the sweep's two multi-lane read patterns (viterbi.cu: load_step, sweep_step): lane = (chain g, destination state j);
// transition row j at j * stride doubles, the chain's exchanged V at g * stride doubles; 128-bit loads of the S values
static void check_sweep_banks(int S, int stride, bool expect_clean)
{
    const int G = 32 / S;
    int worst_lt = 0, worst_x = 0;
    for (int q = 0; 2 * q + 1 < S; q++) {
        int a_lt[32], a_x[32];
        for (int lane = 0; lane < 32; lane++) {
            int g = lane / S;
            const int j = lane - g * S;
            if (g >= G) g = G - 1;
            a_lt[lane] = j * stride * 8 + 16 * q;
            a_x[lane] = g * stride * 8 + 16 * q;
        }
        worst_lt = std::max(worst_lt, wavefronts(a_lt, 16));
        worst_x = std::max(worst_x, wavefronts(a_x, 16));
    }
    if (expect_clean) {
        CHECK(worst_lt == 4);                               // 4 phases, no conflict
        CHECK(worst_x == 4);
    } else {
        CHECK(worst_lt == 16);                              // what ncu measured before the fix
    }
}

// the emission tile of a sweep warp: row r of the TMA box at r * 128 bytes, its 16-byte chunk c at c ^ (r & 7)
// (128-byte swizzle); lane (g, j) reads observation q of row g * S + perm[j].  Returns wavefronts per observation for
// (a) the 8-byte read per step the kernel does today, (b) a 16-byte read of two observations every other step.
static void emission_tile_wavefronts(int S, double* per_step_now, double* per_step_paired)
{
    const int G = 32 / S, normal = S == 3 ? 1 : 2;
    int perm[8], n = 1;
    perm[0] = normal;
    for (int s = 0; s < S; s++) if (s != normal) perm[n++] = s;
    int now = 0, paired = 0;
    for (int q = 0; q < 16; q++) {
        int a8[32], a16[32];
        for (int lane = 0; lane < 32; lane++) {
            int g = lane / S;
            const int j = lane - g * S;
            if (g >= G) g = G - 1;
            const int r = g * S + perm[j];
            a8[lane] = (r * 128 + ((r & 7) << 4)) ^ (q << 3);
            a16[lane] = (r * 128 + ((r & 7) << 4)) ^ ((q & ~1) << 3);
        }
        now += wavefronts(a8, 8);
        if (!(q & 1)) paired += wavefronts(a16, 16);
    }
    *per_step_now = now / 16.0;
    *per_step_paired = paired / 16.0;
}

int main()
{
    {
        // DESIGN.md "what comes next" 1b: today's emission read costs ~5 wavefronts per step at S = 5 (2 would be ideal:
        // rows r and r + 8 of a 16-lane phase share a bank pair); a paired 128-bit read would cost 3
        double now = 0, paired = 0;
        emission_tile_wavefronts(5, &now, &paired);
        CHECK(now >= 4.0 && now <= 6.0);
        CHECK(paired <= 3.0);
        printf("emission tile read at S = 5: %.2f wavefronts per step now, %.2f with paired 128-bit reads\n", now, paired);
    }
    for (int S = 2; S <= 7; S++) check_sweep_banks(S, edb::lt_jstride(S), true);
    check_sweep_banks(7, 8, false);
    for (int S = 2; S <= 7; S++) CHECK(edb::lt_jstride(S) >= S && edb::lt_jstride(S) % 2 == 0 && edb::lt_pitch(S) == S * edb::lt_jstride(S));
    // ---- placement -------------------------------------------------------------------------------------
    const std::vector<int32_t> genome = {19803, 14662, 11433, 11265, 11138, 10912, 10693, 9535, 8919, 8493, 8245, 7930,
                                         7786, 7623, 7476, 6969, 6658, 6490, 6179, 4743, 4167, 3393, 2959, 1984, 595};
    for (int W : {1, 2, 4, 8})
        for (int n_ctas : {1, 11, 137, 148})
            for (int groups : {1, 8, 43, 334}) check_schedule(genome, groups, n_ctas, W);
    for (int W : {1, 2, 4, 8}) check_schedule(std::vector<int32_t>(5, 1002), 128, 148, W);      // the small panel
    check_schedule({7}, 1, 1, 4);
    // ---- CallCNVs transition matrix, column-major T[k + S*j] = P(k -> j) ---------------------------------
    for (int S : {3, 5, 7}) {
        std::vector<double> T(S * S);
        const double tp = 1e-4;
        edb::callcnvs_transitions(S, tp, T.data());
        for (int k = 0; k < S; k++) {
            double row = 0;
            for (int j = 0; j < S; j++) row += T[k + S * j];
            CHECK(std::fabs(row - 1.0) < 1e-15);
        }
        CHECK(T[0] == 1 - tp);
        for (int j = 1; j < S; j++) {
            CHECK(T[0 + S * j] == tp / (S - 1));
            CHECK(T[j + S * 0] == 0.5 && T[j + S * j] == 0.5);
        }
    }
    // ---- framing: as.integer(c(start[1] - 2 L, start, end[last] + 2 L)) ------------------------------------
    {
        const int32_t start[3] = {12012, 13000, 12990}, end[3] = {12057, 13100, 13200};
        int32_t pos[5];
        CHECK(edb::frame_positions(3, start, end, 50000.0, pos) == 0);
        CHECK(pos[0] == 12012 - 100000 && pos[1] == 12012 && pos[2] == 13000 && pos[3] == 12990 && pos[4] == 13200 + 100000);
        const int32_t far_end[1] = {2147483000};
        CHECK(edb::frame_positions(1, start, far_end, 50000.0, pos) != 0);                      // R would give NA
    }
    // ---- log-transition rows against hmm.cpp:62-79 evaluated directly -------------------------------------
    for (int S : {3, 5, 7}) {
        std::vector<double> T(S * S);
        edb::callcnvs_transitions(S, 1e-4, T.data());
        const int32_t pos[6] = {-87988, 12012, 13000, 12990, 12990, 113200};                   // a negative and a zero gap
        for (int pitch : {S * (S + (S & 1)), S * 10}) {
            std::vector<double> lt(6 * pitch, -1.0);
            edb::build_log_transition_rows(S, T.data(), pos, 6, 50000.0, lt.data(), pitch);
            const int js = pitch / S;
            for (int i = 1; i < 6; i++) {
                const double d = std::exp(-(double(pos[i]) - double(pos[i - 1])) / 50000.0);
                for (int j = 0; j < S; j++)
                    for (int k = 0; k < js; k++) {
                        const double got = lt[i * pitch + j * js + k];
                        if (k >= S) { CHECK(got == 0.0); continue; }
                        const double t0 = T[j * S];
                        const double t = k == 0 ? t0 : d * T[j * S + k] + (1.0 - d) * t0;
                        const double want = std::log(t);
                        CHECK((got == want) || (got != got && want != want));                   // same bits, NaN where log(negative)
                    }
            }
        }
    }
    // ---- structure of the CallCNVs table (DESIGN.md "what comes next" 1): per observation and destination state j the
    // row holds at most three distinct values — k = 0, k = j, and one shared by every other source state
    for (int S : {5, 7}) {
        std::vector<double> T(S * S);
        edb::callcnvs_transitions(S, 1e-4, T.data());
        const int32_t pos[5] = {0, 100000, 101685, 101600, 151600};
        const int pitch = S * (S == 7 ? 10 : S + (S & 1)), js = pitch / S;
        std::vector<double> lt(5 * pitch, 0.0);
        edb::build_log_transition_rows(S, T.data(), pos, 5, 50000.0, lt.data(), pitch);
        auto same = [](double a, double b) { return a == b || (a != a && b != b); };
        for (int i = 1; i < 5; i++)
            for (int j = 0; j < S; j++) {
                const double* row = &lt[i * pitch + j * js];
                int other = -1;
                for (int k = 1; k < S; k++) {
                    if (k == j) continue;
                    if (other < 0) other = k;
                    CHECK(same(row[k], row[other]));
                }
            }
    }
    if (!fails) printf("ok\n");
    return fails ? 1 : 0;
}
