// host_tables.cpp — host-side construction of the shared per-bin HMM metadata.
//
// The distance-dependent transition terms of the reference depend on the bin, not on the sample
// (src/hmm.cpp:62-76), yet the reference recomputes 1 exp + 9 log per observation for every sample.
// Here they are computed ONCE per cohort, on the host, with the host libm — the same exp()/log() the
// reference calls — so the table holds the reference's own bits (CUDA's log/exp are <= 1 ulp but not
// bit-identical to glibc, which could flip a near-tie in the backtrace; SURVEY.md §7 hard part 2).
// The table legitimately contains -Inf (log 0) and NaN (log of a negative term when bin starts are
// not monotone, SURVEY.md §8c "NaN edge"); both are kept.  Compiled without FMA contraction.
#include "host_tables.h"
#include "viterbi_step.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <queue>
#include <utility>
#include <cstring>
#include <thread>
#include <vector>

namespace edb {

void callcnvs_transitions(int S, double tp, double* T)
{
    // R/class_definition.R:343-347 (byrow 3x3) generalised as SURVEY.md §8a H4; column-major T[k + S*j]
    for (int k = 0; k < S; k++)
        for (int j = 0; j < S; j++) {
            double p;
            if (k == 0) p = j == 0 ? 1. - tp : tp / (double)(S - 1);
            else p = (j == 0 || j == k) ? 0.5 : 0.0;
            T[k + S * j] = p;
        }
}

static inline uint64_t bits_of(double x)
{
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return u;
}

static void fill_rows(int S, const double* T, const int32_t* pos, double L, int64_t i0, int64_t i1, double* lt, int pitch)
{
    double vals[64], logs[64];
    uint64_t keys[64];
    for (int64_t i = i0; i < i1; i++) {
        const double dist = double(pos[i]) - double(pos[i - 1]);      // hmm.cpp:62
        const double d = std::exp(-dist / L);                         // hmm.cpp:64
        int nuniq = 0;
        double* row = lt + i * pitch;
        const int js = pitch / S;                        // doubles per destination state (S padded to an even count)
        for (int q = 0; q < pitch; q++) row[q] = 0.0;
        for (int j = 0; j < S; j++) {
            const double t0 = T[j * S];
            for (int k = 0; k < S; k++) {
                const double t = k == 0 ? t0 : d * T[j * S + k] + (1.0 - d) * t0;   // hmm.cpp:74-76
                const uint64_t key = bits_of(t);
                int u = 0;
                while (u < nuniq && keys[u] != key) u++;
                if (u == nuniq) {                       // identical inputs give identical log(): memoise per row
                    keys[u] = key;
                    vals[u] = t;
                    logs[u] = std::log(t);              // hmm.cpp:79
                    nuniq++;
                }
                row[j * js + k] = logs[u];
            }
        }
        (void)vals;
    }
}

void build_log_transition_rows(int S, const double* T, const int32_t* pos, int32_t nobs, double L, double* lt, int pitch)
{
    for (int q = 0; q < pitch; q++) lt[q] = 0.0;      // row 0 is never read (hmm.cpp:58 starts at i = 1)
    if (nobs <= 1) return;
    const int64_t n = nobs;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? (hw > 16 ? 16 : hw) : 1);
    if (n < 20000) nt = 1;
    if (nt == 1) { fill_rows(S, T, pos, L, 1, n, lt, pitch); return; }
    std::vector<std::thread> th;
    const int64_t per = (n - 1 + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t a = 1 + t * per, b = a + per < n ? a + per : n;
        if (a >= b) break;
        th.emplace_back(fill_rows, S, T, pos, L, a, b, lt, pitch);
    }
    for (auto& x : th) x.join();
}

void build_decay_rows(const int32_t* pos, int32_t nobs, double L, double* decay)
{
    if (nobs > 0) decay[0] = 0.0;
    for (int32_t i = 1; i < nobs; i++) decay[i] = std::exp(-(double(pos[i]) - double(pos[i - 1])) / L);
}

void nan_to_neg_inf(double* v, size_t n)
{
    for (size_t i = 0; i < n; i++)
        if (v[i] != v[i]) v[i] = -HUGE_VAL;
}

int build_struct_rows(int S, const double* lt, int pitch, int64_t n_rows, StructRow* rows, double* c0, double* c1)
{
    if (S < 3) return 0;
    const int js = pitch / S;
    bool have = false;
    uint64_t k0 = 0, k1 = 0;
    for (int64_t i = 0; i < n_rows; i++) {
        const double* r = lt + i * pitch;
        bool zero = true;
        for (int q = 0; q < pitch && zero; q++) zero = bits_of(r[q]) == 0;
        rows[i] = StructRow{0.0, 0.0, 0.0, 0.0};
        if (zero) continue;
        const uint64_t a0 = bits_of(r[0]), a1 = bits_of(r[js]);              // t(0 -> 0), t(0 -> 1)
        const uint64_t b0 = bits_of(r[1]), sf = bits_of(r[js + 1]), ot = bits_of(r[js + 2]);     // t(k>0 -> 0), t(1 -> 1), t(2 -> 1)
        if (!have) { k0 = a0; k1 = a1; have = true; }
        if (a0 != k0 || a1 != k1) return 0;
        for (int j = 0; j < S; j++)
            for (int k = 0; k < S; k++) {
                const uint64_t v = bits_of(r[j * js + k]);
                const uint64_t want = k == 0 ? (j == 0 ? k0 : k1) : j == 0 ? b0 : k == j ? sf : ot;
                if (v != want) return 0;
            }
        rows[i].b0 = r[1];
        rows[i].sf = r[js + 1];
        rows[i].ot = r[js + 2];                  // t(2 -> 1); S >= 3 so state 2 exists
    }
    if (!have) return 0;
    memcpy(c0, &k0, 8);
    memcpy(c1, &k1, 8);
    // What the speculative step (viterbi_step.h) takes for granted: log-probabilities (<= 0), c0 and c1 finite, and
    // per row  ot - c1 <= b0 - c0  — the "other copy-number state" term loses to k = 0 by at least as much as the
    // return-to-normal term does, so one comparison covers both groups.  With the CallCNVs matrix that is
    // log(1 - d) <= log(1 - d + d/(2(1 - tp))), true for every d; it is verified here, not assumed.
    if (!(std::fabs(*c0) < 1e300) || !(std::fabs(*c1) < 1e300) || *c0 > 0 || *c1 > 0) return 0;
    for (int64_t i = 0; i < n_rows; i++) {
        const StructRow& q = rows[i];
        if (bits_of(q.b0) == 0 && bits_of(q.sf) == 0 && bits_of(q.ot) == 0) continue;          // unused row (see above)
        if (q.b0 > 0 || q.sf > 0 || q.ot > 0 || q.b0 != q.b0 || q.sf != q.sf || q.ot != q.ot) return 0;
        if (q.ot == -HUGE_VAL) continue;
        if (q.b0 == -HUGE_VAL || !(q.ot - *c1 <= q.b0 - *c0 + 9.5367431640625e-07)) return 0;      // 2^-20 of slack
    }
    return 1;
}

void viterbi_schedule(const int32_t* chain_nobs, int n_chains, int groups, int n_ctas, int warps_per_cta,
                      std::vector<int32_t>& begin, std::vector<int32_t>& items)
{
    // sub-partitions that hold sweep warps: 4 per CTA (warp slots w and w + 4 share one), fewer when the CTA has < 4 warps
    const int subs = warps_per_cta < 4 ? warps_per_cta : 4;
    const int n_slots = n_ctas * warps_per_cta, n_parts = n_ctas * subs;
    std::vector<int> order(n_chains);
    for (int c = 0; c < n_chains; c++) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return chain_nobs[x] > chain_nobs[y]; });
    std::vector<int64_t> part_load(n_parts, 0), slot_load(n_slots, 0);
    std::vector<std::vector<int32_t>> per_slot(n_slots);
    // min-heap of (load, partition)
    using Key = std::pair<int64_t, int>;
    std::priority_queue<Key, std::vector<Key>, std::greater<Key>> heap;
    for (int p = 0; p < n_parts; p++) heap.push({0, p});
    for (int oc = 0; oc < n_chains; oc++) {
        const int c = order[oc];
        for (int g = 0; g < groups; g++) {
            Key k = heap.top();
            heap.pop();
            const int p = k.second, cta = p / subs, sub = p % subs;
            int best = -1;
            for (int w = sub; w < warps_per_cta; w += subs) {
                const int s = cta * warps_per_cta + w;
                if (best < 0 || slot_load[s] < slot_load[best]) best = s;
            }
            per_slot[best].push_back(c);
            per_slot[best].push_back(g);
            slot_load[best] += chain_nobs[c];
            part_load[p] += chain_nobs[c];
            heap.push({part_load[p], p});
        }
    }
    begin.assign(n_slots + 1, 0);
    items.clear();
    for (int s = 0; s < n_slots; s++) {
        begin[s] = (int32_t)(items.size() / 2);
        items.insert(items.end(), per_slot[s].begin(), per_slot[s].end());
    }
    begin[n_slots] = (int32_t)(items.size() / 2);
}

// Segmented sweep (viterbi_seam.h): the (chain, 32-sample group) lines of tiles are laid end to end and cut into equal
// shares, one per sweep warp; a line that straddles a share boundary is cut there into pieces, which the kernel sweeps
// concurrently.  A piece that does not start its line costs `warm` warm-up tiles on top of its own.  No piece is
// shorter than min_piece tiles (a cut that would leave a shorter head or tail moves to the line's end), so small
// batches use fewer warps.  Logical share q goes to warp q / n_ctas of CTA q % n_ctas: the shares spread over the SMs
// before they stack up on one.
void viterbi_cut_pieces(const int32_t* chain_tiles, int n_chains, int n_g32, int n_ctas, int warps_per_cta, int warm, int min_piece,
                        std::vector<int32_t>& begin, std::vector<int32_t>& items, std::vector<int32_t>& desc, std::vector<int32_t>& first)
{
    const int n_slots = n_ctas * warps_per_cta;
    if (min_piece < warm + 2) min_piece = warm + 2;
    int64_t total = 0;
    for (int c = 0; c < n_chains; c++) total += (int64_t)chain_tiles[c] * n_g32;
    int64_t target = (total + (int64_t)n_slots * warm + n_slots - 1) / n_slots;
    if (target < min_piece) target = min_piece;
    std::vector<std::vector<int32_t>> per_slot(n_slots);
    desc.clear();
    first.assign((size_t)n_chains * n_g32 + 1, 0);
    int q = 0;
    int64_t room = target;
    for (int c = 0; c < n_chains; c++)
        for (int g = 0; g < n_g32; g++) {
            first[(size_t)c * n_g32 + g] = (int32_t)(desc.size() / 4);
            const int nt = chain_tiles[c];
            int pos = 0;
            while (pos < nt) {
                const int cost0 = pos > 0 ? warm : 0;
                int64_t take = room - cost0;
                if (q == n_slots - 1) take = nt - pos;                       // the last share takes what is left
                if (take < min_piece) take = min_piece;
                if (take > nt - pos) take = nt - pos;
                if (nt - (pos + take) < min_piece) take = nt - pos;          // no short tail
                const int piece = (int)(desc.size() / 4);
                desc.push_back(c);
                desc.push_back(g);
                desc.push_back(pos);
                desc.push_back(pos + (int)take);
                const int slot = (q % n_ctas) * warps_per_cta + q / n_ctas;
                per_slot[slot].push_back(piece);
                per_slot[slot].push_back(0);
                room -= take + cost0;
                pos += (int)take;
                if (room <= 0 && q < n_slots - 1) {
                    q++;
                    room += target;
                    if (room < min_piece) room = min_piece;
                }
            }
        }
    first[(size_t)n_chains * n_g32] = (int32_t)(desc.size() / 4);
    begin.assign(n_slots + 1, 0);
    items.clear();
    for (int s = 0; s < n_slots; s++) {
        begin[s] = (int32_t)(items.size() / 2);
        items.insert(items.end(), per_slot[s].begin(), per_slot[s].end());
    }
    begin[n_slots] = (int32_t)(items.size() / 2);
}

int64_t pack_counts(int bits, const int32_t* counts, int64_t stride, int32_t n_samples, int64_t n_bins, void* out, int64_t out_stride,
                    int64_t* ovf_index, int32_t* ovf_value, int64_t cap)
{
    struct Entry { int64_t index; int32_t value; };
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<int64_t>(hw ? (hw > 16 ? 16 : hw) : 1, n_samples);
    if ((int64_t)n_samples * n_bins < (1 << 20)) nt = 1;
    if (nt < 1) nt = 1;
    std::vector<std::vector<Entry>> found(nt);
    std::vector<int> bad(nt, 0);
    const int32_t sentinel = bits == 12 ? 4095 : 65535;
    // a block of consecutive samples per thread: the per-thread lists, concatenated in thread order, are sorted by flat index
    auto work = [&](int t) {
        const int32_t s0 = (int32_t)((int64_t)n_samples * t / nt), s1 = (int32_t)((int64_t)n_samples * (t + 1) / nt);
        for (int32_t s = s0; s < s1; s++) {
            const int32_t* row = counts + (int64_t)s * stride;
            auto field = [&](int64_t b) -> uint32_t {
                const int32_t v = row[b];
                if (v < 0) bad[t] = 1;
                if (v >= sentinel) {
                    found[t].push_back(Entry{(int64_t)s * n_bins + b, v});
                    return (uint32_t)sentinel;
                }
                return (uint32_t)v;
            };
            if (bits == 16) {
                uint16_t* o = (uint16_t*)out + (int64_t)s * out_stride;
                for (int64_t b = 0; b < n_bins; b++) o[b] = (uint16_t)field(b);
            } else {
                uint8_t* o = (uint8_t*)out + (int64_t)s * out_stride;
                int64_t b = 0;
                for (; b + 1 < n_bins; b += 2, o += 3) {
                    const uint32_t w = field(b) | field(b + 1) << 12;             // bins 2i, 2i+1 in bytes 3i .. 3i+2
                    o[0] = (uint8_t)w;
                    o[1] = (uint8_t)(w >> 8);
                    o[2] = (uint8_t)(w >> 16);
                }
                if (b < n_bins) {
                    const uint32_t w = field(b);
                    o[0] = (uint8_t)w;
                    o[1] = (uint8_t)(w >> 8);
                    o[2] = 0;
                }
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    int64_t n = 0;
    for (int t = 0; t < nt; t++) {
        if (bad[t]) return -1;
        for (const Entry& e : found[t]) {
            if (n < cap) {
                ovf_index[n] = e.index;
                ovf_value[n] = e.value;
            }
            n++;
        }
    }
    return n;
}

int frame_positions(int64_t nb, const int32_t* start, const int32_t* end, double L, int32_t* pos)
{
    // R/class_definition.R:368: as.integer(c(start[1] - 2*L, start, end[last] + 2*L))
    const double head = (double)start[0] - 2 * L, tail = (double)end[nb - 1] + 2 * L;
    if (!(std::fabs(head) < 2147483647.0) || !(std::fabs(tail) < 2147483647.0)) return 1;   // R would give NA
    pos[0] = (int32_t)head;                       // as.integer truncates toward zero
    for (int64_t b = 0; b < nb; b++) pos[b + 1] = start[b];
    pos[nb + 1] = (int32_t)tail;
    return 0;
}

}  // namespace edb
