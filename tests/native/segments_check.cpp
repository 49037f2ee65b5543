// CPU check of the segmented Viterbi sweep's host-visible logic, compiled and run by tests/test_host_tables.py:
//  (1) host_tables.cpp: viterbi_cut_pieces — every tile of every (chain, 32-sample group) line is covered exactly once,
//      a line's pieces are consecutive, no piece is shorter than asked, the shares are balanced;
//  (2) viterbi_step.h: viterbi_step_margin against the plain scan of src/hmm.cpp:66-88 (values and first maximum, bit for
//      bit) and its list of low-lead decisions against leads computed here; viterbi_step_spec_m against viterbi_step_spec;
//  (3) the whole scheme of viterbi_seam.h on the host: chains of a few thousand observations with reference-sized
//      magnitudes (|V| up to 1e7) are swept sequentially (the reference's arithmetic) and as pieces from warm-ups, the
//      pieces' decisions are certified with seam_advance and the error multipliers exactly as the kernel does it, and
//      every certified decision must equal the sequential sweep's; CNV regions across seams, uninformative stretches
//      longer than the warm-up (seams that must be refused), near-ties planted at the listed threshold.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "host_tables.h"
#include "viterbi_seam.h"
#include "viterbi_step.h"

static int fails = 0;
#define CHECK(c, ...)                     \
    do {                                  \
        if (!(c)) {                       \
            if (fails < 20) {             \
                printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                printf(__VA_ARGS__);      \
                printf("\n");             \
            }                             \
            fails++;                      \
        }                                 \
    } while (0)

static uint64_t bits(double x)
{
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

// ---------------------------------------------------------------------------------------------- (1)
static void check_cut(const std::vector<int32_t>& tiles, int n_g32, int n_ctas, int W, int warm, int min_piece, double max_imbalance)
{
    std::vector<int32_t> begin, items, desc, first;
    edb::viterbi_cut_pieces(tiles.data(), (int)tiles.size(), n_g32, n_ctas, W, warm, min_piece, begin, items, desc, first);
    const int n_pieces = (int)desc.size() / 4, n_lines = (int)tiles.size() * n_g32;
    CHECK((int)first.size() == n_lines + 1 && first[n_lines] == n_pieces, "first[] size");
    const int mp = std::max(min_piece, warm + 2);
    for (int l = 0; l < n_lines; l++) {
        const int c = l / n_g32, g = l % n_g32;
        int pos = 0;
        for (int p = first[l]; p < first[l + 1]; p++) {
            CHECK(desc[4 * p] == c && desc[4 * p + 1] == g, "piece %d belongs to another line", p);
            CHECK(desc[4 * p + 2] == pos, "piece %d starts at %d, expected %d", p, desc[4 * p + 2], pos);
            const int len = desc[4 * p + 3] - desc[4 * p + 2];
            CHECK(len >= mp || (first[l + 1] - first[l] == 1), "piece %d of line %d has %d tiles (< %d)", p, l, len, mp);
            CHECK(len > 0, "empty piece");
            if (pos > 0) CHECK(pos - warm >= 1, "warm-up of piece %d would start before tile 1", p);
            pos = desc[4 * p + 3];
        }
        CHECK(pos == tiles[c], "line %d covered up to %d of %d", l, pos, tiles[c]);
        CHECK(tiles[c] > 0 || first[l + 1] == first[l], "pieces for an empty chain");
    }
    // every piece dealt exactly once; loads
    std::vector<int> seen(n_pieces, 0);
    const int n_slots = n_ctas * W;
    CHECK((int)begin.size() == n_slots + 1, "begin size");
    int64_t total = 0, busiest = 0;
    int used = 0;
    for (int s = 0; s < n_slots; s++) {
        int64_t load = 0;
        for (int q = begin[s]; q < begin[s + 1]; q++) {
            const int p = items[2 * q];
            CHECK(p >= 0 && p < n_pieces, "piece id");
            seen[p]++;
            load += desc[4 * p + 3] - desc[4 * p + 2] + (desc[4 * p + 2] > 0 ? warm : 0);
        }
        total += load;
        busiest = std::max(busiest, load);
        used += load > 0;
    }
    for (int p = 0; p < n_pieces; p++) CHECK(seen[p] == 1, "piece %d dealt %d times", p, seen[p]);
    if (max_imbalance > 0 && used > 0) {
        const double avg = (double)total / used;
        CHECK(busiest <= max_imbalance * avg + mp, "imbalance: busiest %lld, average %.1f over %d warps", (long long)busiest, avg, used);
    }
    printf("cut: %zu chains x %d groups, %d pieces on %d of %d warps, busiest %lld tiles (average %.1f)\n", tiles.size(), n_g32, n_pieces, used,
           n_slots, (long long)busiest, used ? (double)total / used : 0.0);
}

// ---------------------------------------------------------------------------------------------- (2), (3)
template <int S>
struct Model {
    double c0, c1;
    std::vector<edb::StructRow> rows;
    std::vector<double> em;         // [i * S + j]
    int n;
};

template <int S>
static Model<S> make_model(uint64_t seed, int n, double scale, int flat_from, int flat_len, double near_ties)
{
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    Model<S> m;
    m.n = n;
    const double tp = 1e-4;
    m.c0 = std::log(1 - tp);
    m.c1 = std::log(tp / (S - 1));
    m.rows.resize(n);
    m.em.resize((size_t)n * S);
    int cnv_left = 0, cnv_state = 0;
    for (int i = 0; i < n; i++) {
        const double d = std::exp(-U(rng) * 3);
        m.rows[i].b0 = std::log(d * 0.5 + (1.0 - d) * (1 - tp));
        m.rows[i].sf = std::log(d * 0.5 + (1.0 - d) * (tp / (S - 1)));
        m.rows[i].ot = std::log((1.0 - d) * (tp / (S - 1)));
        m.rows[i].pad = 0;
    }
    for (int i = 0; i < n; i++) {
        if (cnv_left == 0 && U(rng) < 0.004) {
            cnv_left = 1 + (int)(rng() % 60);       // regions up to 60 observations: longer than a 32-observation warm-up
            cnv_state = 1 + (int)(rng() % (S - 1));
        }
        const double base = -scale * (0.2 + U(rng));
        for (int j = 0; j < S; j++) {
            double llr = -(2.0 + 30.0 * U(rng));    // CNV states lose
            if (U(rng) < 0.08) llr = 0.5 + 3.0 * U(rng);   // noisy bin: a CNV state beats normal for one observation
            if (cnv_left > 0 && j == cnv_state) llr = 3.0 + 20.0 * U(rng);
            // after a step won from k = 0, V[j] - V[0] = llr + c1 - c0, so at the NEXT observation self - cand0 = llr + sf' - c0:
            // this llr puts that lead within near_ties / 2 of zero (listed when below 2^-14)
            if (near_ties > 0 && i + 1 < n && U(rng) < 0.01) llr = m.c0 - m.rows[i + 1].sf + (U(rng) - 0.5) * near_ties;
            m.em[(size_t)i * S + j] = j == 0 ? base : base + llr;
        }
        if (i >= flat_from && i < flat_from + flat_len)
            for (int j = 0; j < S; j++) m.em[(size_t)i * S + j] = 0.0;          // no reads: every state's likelihood is exactly 0
        if (cnv_left > 0) cnv_left--;
    }
    return m;
}

template <int S>
static void expand_row(double c0, double c1, const edb::StructRow& r, double* lt)
{
    for (int j = 0; j < S; j++)
        for (int k = 0; k < S; k++) lt[j * S + k] = k == 0 ? (j == 0 ? c0 : c1) : j == 0 ? r.b0 : k == j ? r.sf : r.ot;
}

template <int S>
static void check_steps(uint64_t seed)
{
    Model<S> m = make_model<S>(seed, 4000, 400.0, -1, 0, 4e-4);
    double Va[S], Vb[S], Vc[S], Vd[S];
    for (int j = 0; j < S; j++) Va[j] = Vb[j] = j == 0 ? 0.0 : -1.0 - j;
    long long close_seen = 0, spec_ok = 0;
    for (int i = 0; i < m.n; i++) {
        double lt[S * S];
        expand_row<S>(m.c0, m.c1, m.rows[i], lt);
        unsigned aa[S], ab[S];
        // leads by hand, from the V of before the step
        unsigned want_close = 0;
        for (int j = 0; j < S; j++) {
            double c[S], best = -HUGE_VAL, second = -HUGE_VAL;
            int bi = -1;
            for (int k = 0; k < S; k++) {
                c[k] = (m.em[(size_t)i * S + j] + Va[k]) + lt[j * S + k];
                if (c[k] > best) best = c[k], bi = k;
            }
            for (int k = 0; k < S; k++)
                if (k != bi && c[k] > second) second = c[k];
            if (!(best - second >= edb::kSegTau)) want_close |= 1u << j;
        }
        for (int j = 0; j < S; j++) Vc[j] = Vd[j] = Va[j];
        edb::viterbi_step_scan<S>(Va, &m.em[(size_t)i * S], lt, aa);
        const unsigned close = edb::viterbi_step_margin<S>(Vb, &m.em[(size_t)i * S], m.c0, m.c1, m.rows[i], ab);
        for (int j = 0; j < S; j++) {
            CHECK(bits(Va[j]) == bits(Vb[j]), "margin step value, S=%d step %d state %d", S, i, j);
            CHECK(aa[j] == ab[j], "margin step winner, S=%d step %d state %d: %u vs %u", S, i, j, aa[j], ab[j]);
        }
        CHECK(close == want_close, "listed decisions, S=%d step %d: %x vs %x", S, i, close, want_close);
        close_seen += close != 0;
        // speculative step with leads vs the one without: same values, bits and acceptance; min_hi = smallest |self - cand0|
        bool ok1 = true, ok2 = true;
        unsigned mh = 0x7FFFFFFFu;
        const double c0m = m.c0 - edb::kSpecMargin;
        const unsigned b1 = edb::viterbi_step_spec<S>(Vc, &m.em[(size_t)i * S], m.c0, m.c1, c0m, m.rows[i], ok1);
        const unsigned b2 = edb::viterbi_step_spec_m<S>(Vd, &m.em[(size_t)i * S], m.c0, m.c1, c0m, m.rows[i], ok2, mh);
        CHECK(ok1 == ok2 && b1 == b2, "spec_m acceptance / bits");
        for (int j = 0; j < S; j++) CHECK(bits(Vc[j]) == bits(Vd[j]), "spec_m values");
        if (ok1) {
            spec_ok++;
            for (int j = 0; j < S; j++) CHECK(bits(Vc[j]) == bits(Va[j]), "accepted speculative step differs from the scan");
            // an accepted step with min_hi >= kSegTauHi must have no listed decision among destinations > 0 ... and none at 0
            if (mh >= edb::kSegTauHi) CHECK(close == 0, "speculative step accepted with leads >= 2^-14, yet the scan lists %x (step %d)", close, i);
        }
    }
    printf("steps S=%d: %lld of %d with a listed decision, %lld accepted speculatively\n", S, close_seen, m.n, spec_ok);
}

// sequential sweep (reference arithmetic) vs pieces + certification, as the kernel and the check kernel do it
template <int S>
static void check_scheme(uint64_t seed, int n, double scale, int piece, int warm, int flat_from, int flat_len, double near_ties, bool expect_uncertified)
{
    Model<S> m = make_model<S>(seed, n, scale, flat_from, flat_len, near_ties);
    // reference
    std::vector<unsigned> ref_arg((size_t)n * S);
    std::vector<double> ref_v((size_t)n * S);
    double V[S];
    for (int j = 0; j < S; j++) V[j] = j == 0 ? 0.0 : -HUGE_VAL;
    for (int i = 0; i < n; i++) {
        double lt[S * S];
        expand_row<S>(m.c0, m.c1, m.rows[i], lt);
        unsigned a[S];
        edb::viterbi_step_scan<S>(V, &m.em[(size_t)i * S], lt, a);
        for (int j = 0; j < S; j++) ref_arg[(size_t)i * S + j] = a[j], ref_v[(size_t)i * S + j] = V[j];
    }
    // reference path (ends in state 0)
    std::vector<int> path(n);
    {
        int st = 0;
        for (int i = n - 1; i >= 0; i--) {
            path[i] = st;
            st = (int)ref_arg[(size_t)i * S + st];
            if (st == 7) st = 0;
        }
    }
    // pieces: sweep them all (they are independent), then judge every seam on its own with the line's bound of |C|
    struct Piece {
        int p0, p1;
        double x_in[S], x_end[S];
        edb::PieceErr pe;
        std::vector<unsigned> arg, closev;
        std::vector<float> lead;        // per decision, as the kernel lists it: a float rounded towards zero, low three bits dropped
        double certified = 0.0;         // the lead above which the piece's decisions are certified (seam_check)
    };
    std::vector<Piece> pieces;
    long long certified = 0, listed = 0, listed_on_path = 0, listed_certified = 0, wrong_certified = 0, refused_pieces = 0;
    const double c0m = m.c0 - edb::kSpecMargin;
    double cabs = 0.0;
    for (int p0 = 0; p0 < n; p0 += piece) {
        Piece pc;
        pc.p0 = p0;
        pc.p1 = std::min(n, p0 + piece);
        const int p1 = pc.p1;
        const bool mseg = p0 > 0;
        double X[S];
        for (int j = 0; j < S; j++) X[j] = j == 0 ? 0.0 : -HUGE_VAL;
        for (int i = mseg ? p0 - warm : 0; i < p0; i++) {                  // warm-up: not recorded
            unsigned a[S];
            double lt[S * S];
            expand_row<S>(m.c0, m.c1, m.rows[i], lt);
            edb::viterbi_step_scan<S>(X, &m.em[(size_t)i * S], lt, a);
        }
        for (int j = 0; j < S; j++) pc.x_in[j] = X[j];
        edb::PieceErr pe{0, 0, mseg ? 1u : 0u, 0, 0, 0};
        unsigned err_a = mseg ? 1u : 0u, err_b = 0;
        pc.arg.assign((size_t)(p1 - p0) * S, 0);
        pc.closev.assign(p1 - p0, 0);
        pc.lead.assign((size_t)(p1 - p0) * S, 0.0f);
        for (int i = p0; i < p1; i++) {
            for (int j = 0; j < S; j++) {
                pe.mag_v = std::max(pe.mag_v, edb::f64_hi(X[j]) << 1);
                pe.mag_e = std::max(pe.mag_e, edb::f64_hi(m.em[(size_t)i * S + j]) << 1);
            }
            unsigned a[S];
            if (!mseg) {
                double lt[S * S];
                expand_row<S>(m.c0, m.c1, m.rows[i], lt);
                edb::viterbi_step_scan<S>(X, &m.em[(size_t)i * S], lt, a);
            } else {
                // the kernel's choice: speculative step when accepted with every lead >= 2^-14, else the scan with runner-up
                double Xs[S];
                for (int j = 0; j < S; j++) Xs[j] = X[j];
                bool ok = (edb::f64_hi(X[0]) << 1) < edb::kSpecBigHi2;
                unsigned mh = 0x7FFFFFFFu;
                const unsigned b = edb::viterbi_step_spec_m<S>(Xs, &m.em[(size_t)i * S], m.c0, m.c1, c0m, m.rows[i], ok, mh);
                if (ok && mh >= edb::kSegTauHi) {
                    for (int j = 0; j < S; j++) X[j] = Xs[j];
                    a[0] = 0;
                    for (int j = 1; j < S; j++) a[j] = (b >> (j - 1) & 1u) ? (unsigned)j : 0u;
                    edb::seg_err_step(err_a, err_b, b ? 1 : 0);
                } else {
                    double ld[S];
                    const unsigned close = edb::viterbi_step_margin<S>(X, &m.em[(size_t)i * S], m.c0, m.c1, m.rows[i], a, ld);
                    pc.closev[i - p0] = close;
                    for (int j = 0; j < S; j++) {
                        float f = (float)(ld[j] > 0.0 ? ld[j] : 0.0);
                        if ((double)f > ld[j]) f = std::nextafterf(f, 0.0f);           // round towards zero
                        uint32_t u;
                        memcpy(&u, &f, 4);
                        u &= ~7u;
                        memcpy(&f, &u, 4);
                        pc.lead[(size_t)(i - p0) * S + j] = f;
                    }
                    edb::seg_err_step(err_a, err_b, edb::seg_err_kind<S>(a, close));
                }
                pe.max_b = std::max(pe.max_b, err_b);
            }
            for (int j = 0; j < S; j++) pc.arg[(size_t)(i - p0) * S + j] = a[j];
        }
        for (int j = 0; j < S; j++) {
            pe.mag_v = std::max(pe.mag_v, edb::f64_hi(X[j]) << 1);
            pc.x_end[j] = X[j];
        }
        pe.end_a = err_a;
        pe.end_b = err_b;
        pc.pe = pe;
        cabs += edb::piece_cabs_share(mseg ? pe.mag_v : edb::f64_hi(X[0]) << 1, p1 - p0);      // (as the kernel: the exact piece through its last V[0])
        pieces.push_back(std::move(pc));
    }
    // the bound of |C| must hold: compare with the real C = R[0] - X[0] at every piece's end
    {
        size_t q = 0;
        for (const Piece& pc : pieces) {
            const double C = ref_v[(size_t)(pc.p1 - 1) * S] - pc.x_end[0];
            CHECK(std::fabs(C) <= cabs, "piece %zu: |C| = %g exceeds the line's bound %g", q, std::fabs(C), cabs);
            q++;
        }
    }
    std::vector<int> refused(pieces.size(), 0);
    for (size_t q = 1; q < pieces.size(); q++) {
        Piece& pc = pieces[q];
        const int bad = edb::seam_check<S>(pc.x_in, pieces[q - 1].x_end, pc.pe, q == 1 ? nullptr : &pieces[q - 1].pe, cabs, &pc.certified);
        if (bad && getenv("SEG_DEBUG"))
            printf("  seam %zu bad %d: cabs %g mag_v %x mag_e %x max_b %u end_a %u end_b %u prev(end_a %u end_b %u) rho %g\n", q, bad, cabs, pc.pe.mag_v, pc.pe.mag_e,
                   pc.pe.max_b, pc.pe.end_a, pc.pe.end_b, pieces[q - 1].pe.end_a, pieces[q - 1].pe.end_b, edb::piece_rho(pc.pe, cabs));
        refused[q] = bad != 0;
        refused_pieces += bad != 0;
    }
    // (the kernel sends a chain with any refused seam to the repair pass; here the pieces are judged one by one, a piece
    // counting as certified only if every seam up to it closed — the induction of viterbi_seam.h)
    bool chain_ok = true;
    for (size_t q = 0; q < pieces.size(); q++) {
        chain_ok = chain_ok && !refused[q];
        if (!chain_ok) break;
        const Piece& pc = pieces[q];
        for (int i = pc.p0; i < pc.p1; i++)
            for (int j = 0; j < S; j++) {
                const bool is_listed = pc.closev[i - pc.p0] >> j & 1u;
                // a listed decision is certified after all when its lead exceeds what the piece's deviation reached
                if (is_listed && !((double)pc.lead[(size_t)(i - pc.p0) * S + j] > pc.certified)) {
                    listed++;
                    listed_on_path += path[i] == j;
                    continue;
                }
                listed_certified += is_listed;
                certified++;
                if (pc.arg[(size_t)(i - pc.p0) * S + j] != ref_arg[(size_t)i * S + j]) wrong_certified++;
            }
    }
    CHECK(wrong_certified == 0, "S=%d seed %llu: %lld certified decisions differ from the sequential sweep", S, (unsigned long long)seed, wrong_certified);
    CHECK(certified > 0, "nothing certified");
    CHECK(refused_pieces == 0, "S=%d seed %llu scale %.0f: %lld pieces refused", S, (unsigned long long)seed, scale, refused_pieces);
    if (expect_uncertified) CHECK(listed > 0, "S=%d: leads planted inside the rounding noise were all certified", S);
    printf("scheme S=%d scale %.0f piece %d warm %d: certified %lld (%lld of them listed, lead above the piece's bound), uncertified %lld (%lld on the path), "
           "refused pieces %lld\n", S, scale, piece, warm, certified, listed_certified, listed, listed_on_path, refused_pieces);
}

int main()
{
    // (1) the bench geometry (25 chromosomes, 12,516 tiles) at 256 and 2,000 samples, small and degenerate shapes
    const std::vector<int32_t> hg = {1238, 917, 715, 705, 697, 640, 600, 580, 560, 540, 520, 500, 480, 460, 440, 420, 400, 380, 360, 340, 320, 300, 280, 124, 0};
    check_cut(hg, 8, 148, 4, 4, 32, 1.10);
    check_cut(hg, 63, 148, 4, 4, 32, 1.10);
    check_cut({20, 20, 20, 20, 1, 0, 3}, 1, 148, 4, 1, 3, 0);
    check_cut({400}, 3, 2, 4, 2, 6, 1.5);
    check_cut({7, 90, 33}, 2, 148, 4, 4, 32, 0);
    std::mt19937_64 rng(7);
    for (int t = 0; t < 200; t++) {
        std::vector<int32_t> tiles(1 + rng() % 30);
        for (auto& x : tiles) x = (int32_t)(rng() % 1500);
        check_cut(tiles, 1 + (int)(rng() % 70), 1 + (int)(rng() % 148), 4, 1 + (int)(rng() % 8), (int)(rng() % 64), 0);
    }
    // (2)
    check_steps<3>(11);
    check_steps<5>(12);
    check_steps<7>(13);
    // (3) magnitudes of the reference's likelihood (no binomial coefficient: hundreds per bin, |V| ~ 1e7 per chromosome)
    for (uint64_t seed = 1; seed <= 6; seed++) {
        check_scheme<5>(seed, 12000, 650.0, 1500, 64, -1, 0, 0.0, false);
        check_scheme<5>(seed + 10, 12000, 5.0, 700, 32, -1, 0, 4e-4, false);
        check_scheme<3>(seed + 20, 8000, 650.0, 333, 32, -1, 0, 4e-4, false);
        // leads planted INSIDE the rounding noise of |V| ~ 1e7 (one ulp is 2e-9): they stay uncertified — and may well differ from
        // the sequential sweep, which is what the repair pass is for — while everything certified must still agree
        check_scheme<5>(seed + 40, 12000, 650.0, 1500, 32, -1, 0, 2e-8, true);
        check_scheme<7>(seed + 30, 8000, 100.0, 1000, 64, -1, 0, 0.0, false);
    }
    // an uninformative stretch that swallows a whole warm-up: the piece starts from (0, -Inf, ...) one observation before...
    // no — from the stationary vector of the flat stretch, which the previous piece also reaches: certified; and a warm-up
    // that starts INSIDE a called region which extends to the seam: refused or certified, never wrong
    check_scheme<5>(99, 6000, 650.0, 1000, 16, 900, 200, 0.0, false);
    if (fails) {
        printf("%d failures\n", fails);
        return 1;
    }
    printf("ok\n");
    return 0;
}
