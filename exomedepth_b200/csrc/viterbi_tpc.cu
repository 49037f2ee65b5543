// viterbi_tpc.cu — Viterbi sweep with ONE THREAD PER CHAIN, for CallCNVs-structured transition rows (sm_100a).
//
// Same recurrence, same bits as src/hmm.cpp:58-90 (see viterbi_step.h for why the structured step is exact).  Where
// viterbi.cu spreads a chain's S destination states over S lanes and pays a shared-memory exchange of V on every one
// of the ~20,000 dependent steps of chromosome 1 (~150 cycles per step however few warps share the SM), here a lane
// keeps all S values of V in registers: no exchange, 32 chains per warp instead of 32/S, and per step ~45 FP64
// instructions instead of 14 per lane x S lanes.  A warp = the same chromosome of 32 consecutive samples, so the
// distance-dependent transition terms of an observation are warp-uniform (three doubles, broadcast loads).
//
// Warp roles as in viterbi.cu: W sweep warps and one producer warp per CTA; lane w of the producer feeds sweep warp
// w's ring: per 8-observation half-tile one 2-D TMA load of the emissions (box 8 bins x 32*S rows, 64-byte swizzle:
// a quarter-warp's 128-bit reads of two observations hit 8 distinct 16-byte chunks) and one bulk copy of the half's 8
// StructRows, completing on the stage's `full` mbarrier; the sweep warp hands the stage back through `empty`.
// Back-pointers leave in the record layout of viterbi_common.cuh, so tilemap / trace / expand are shared.
#include <algorithm>
#include <cuda.h>

#include "host_tables.h"
#include "kernels.cuh"
#include "tma_ptx.cuh"
#include "viterbi_common.cuh"
#include "viterbi_seam.h"
#include "viterbi_step.h"

namespace edb {

// A ring stage holds HALF a record: 8 observations of the warp's 32 chains (32*S rows x 64 bytes, 64-byte swizzle)
// and their 8 StructRows — 10.5 KB at S = 5, so that four sweep warps (one per SM sub-partition) get five stages each.
// With whole 16-observation tiles (21 KB) four warps got two stages and the sweep stalled on its loads (2.8 vs 2.05 ms
// at 256 x 200k x 5), two warps per CTA got five but took twice the SMs.
constexpr int kHalf = kTile / 2;
__host__ __device__ constexpr int tpc_em_bytes(int S) { return 32 * S * kHalf * 8; }                 // 32 chains x S rows x 64 B
__host__ __device__ constexpr int tpc_stage_bytes(int S) { return (tpc_em_bytes(S) + kHalf * (int)sizeof(StructRow) + 511) / 512 * 512; }
__host__ __device__ constexpr int tpc_stages(int S, int W)
{
    const int fit = (214 * 1024) / (W * tpc_stage_bytes(S));
    return fit > 10 ? 10 : fit;
}

// ---- the exact pair: out of line, taken when the speculative steps of a pair of observations do not apply --------
// (a group of copy-number states may win or tie — inside and around CNV regions, a few percent of the pairs of a
// warp — or an emission is NaN / +-Inf / huge: pathological phi).  By-value arguments and result: the caller's V
// stays in registers.
template <int S>
struct PairArgs {
    double v[S];
    uint32_t em[S];             // shared-memory address of this lane's two emissions (observations A, B) per HMM state
    uint32_t rows;              // shared-memory address of the two StructRows
    double c0, c1;
    // segmented sweep only (tpc_pair_margin): where to report, and which decisions these are
    const ViterbiArgs* va;
    int chain, piece, smp, obs; // obs: the first observation of the pair
    int rec;                    // 0: warm-up (decisions are not recorded, hence not listed)
};
template <int S>
struct PairRes {
    double v[S];
    unsigned acc[S];            // per destination: back-pointer of the first observation | second << 4
    unsigned kinds;             // tpc_pair_margin: how the two steps move the error multipliers (seg_err_step), bit 0 / bit 1
};

template <int S>
__device__ __noinline__ PairRes<S> tpc_pair_exact(const PairArgs<S> a)
{
    PairRes<S> r;
    double V[S], ea[S], eb[S];
    unsigned worst = (unsigned)__double2hiint(a.v[0]) << 1;
#pragma unroll
    for (int j = 0; j < S; j++) {
        V[j] = a.v[j];
        const double2 e = lds_f64x2(a.em[j]);
        ea[j] = e.x;
        eb[j] = e.y;
        worst = max(worst, max((unsigned)__double2hiint(e.x) << 1, (unsigned)__double2hiint(e.y) << 1));
    }
    const double2 r0a = lds_f64x2(a.rows), r1a = lds_f64x2(a.rows + 32);
    const StructRow rowA{r0a.x, r0a.y, lds_f64(a.rows + 16), 0.0}, rowB{r1a.x, r1a.y, lds_f64(a.rows + 48), 0.0};
    if (worst >= 0xFFE00000u) {                             // NaN / Inf in sight: the fully checked step
        unsigned arg[S];
        viterbi_step_struct<S>(V, ea, a.c0, a.c1, rowA, r.acc);
        viterbi_step_struct<S>(V, eb, a.c0, a.c1, rowB, arg);
#pragma unroll
        for (int j = 0; j < S; j++) r.acc[j] |= arg[j] << 4;
    } else {
        double VpB[S];
        unsigned argB[S];
        const unsigned needA = viterbi_step_fast<S>(V, ea, a.c0, a.c1, rowA, r.acc);
#pragma unroll
        for (int j = 0; j < S; j++) VpB[j] = V[j];
        const unsigned needB = viterbi_step_fast<S>(V, eb, a.c0, a.c1, rowB, argB);
#pragma unroll
        for (int j = 0; j < S; j++) {
            r.acc[j] |= argB[j] << 4;
            if (needA >> j & 1u) r.acc[j] = (r.acc[j] & 0xF0u) | viterbi_resolve_arg<S>(a.v, ea[j], j, a.c0, a.c1, rowA);
            if (needB >> j & 1u) r.acc[j] = (r.acc[j] & 0x0Fu) | (viterbi_resolve_arg<S>(VpB, eb[j], j, a.c0, a.c1, rowB) << 4);
        }
    }
#pragma unroll
    for (int j = 0; j < S; j++) r.v[j] = V[j];
    return r;
}

// ---- segmented sweep: flags and the certifying step -----------------------------------------------------------------
__device__ __forceinline__ void seg_mark_bad(int32_t* flags, int n_chains, int n_samples, int chain, int smp, int why)
{
    flags[1] = 1;
    flags[seg_off_chain(chain)] = 1;
    flags[seg_off_line(n_chains, n_samples, chain, smp >> 5)] = 1;
    atomicOr(&flags[seg_off_pair(n_chains, n_samples, chain, smp)], why);
}

// One observation of a piece that did not start its chain, whenever the speculative step does not apply: the plain
// scan with the runner-up (viterbi_step.h: viterbi_step_margin); decisions with a lead below kSegTau are listed for the
// check kernel.  A NaN / Inf in sight — or more listed decisions than the list holds — sends the chain to the repair pass.
// Returns the step's kind for the error multipliers.
template <int S>
__device__ __forceinline__ int seg_step_margin(double* V, const double* em, const StructRow& row, double c0, double c1, unsigned* arg,
                                               const ViterbiArgs* va, int chain, int piece, int smp, int obs, int rec)
{
    unsigned worst = 0;
#pragma unroll
    for (int j = 0; j < S; j++) worst = max(worst, max((unsigned)__double2hiint(V[j]) << 1, (unsigned)__double2hiint(em[j]) << 1));
    if (worst >= 0xFFE00000u) {
        viterbi_step_struct<S>(V, em, c0, c1, row, arg);
        if (rec) seg_mark_bad(va->seg_flags, va->n_chains, va->n_samples, chain, smp, kBadNonFinite);
        return 1;
    }
    double lead[S];
    const unsigned close = viterbi_step_margin<S>(V, em, c0, c1, row, arg, lead);
    if (close && rec) {
#pragma unroll 1
        for (int j = 0; j < S; j++)
            if (close >> j & 1u) {
                // (piece, sample, observation, destination | the lead as a float rounded towards zero, low three bits dropped)
                const double ld = lead[j] > 0.0 ? lead[j] : 0.0;
                const int q = atomicAdd(va->seg_flags, 1);
                if (q < va->seg_close_cap) va->seg_close[q] = make_int4(piece, smp, obs, (int)((__float_as_uint(__double2float_rz(ld)) & ~7u) | (unsigned)j));
                else seg_mark_bad(va->seg_flags, va->n_chains, va->n_samples, chain, smp, kBadListFull);
            }
    }
    return seg_err_kind<S>(arg, close);
}

// the pair of observations the speculative steps did not settle, for a piece that did not start its chain (out of line)
template <int S>
__device__ __noinline__ PairRes<S> tpc_pair_margin(const PairArgs<S> a)
{
    PairRes<S> r;
    double V[S], ea[S], eb[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        V[j] = a.v[j];
        const double2 e = lds_f64x2(a.em[j]);
        ea[j] = e.x;
        eb[j] = e.y;
    }
    const double2 r0a = lds_f64x2(a.rows), r1a = lds_f64x2(a.rows + 32);
    const StructRow rowA{r0a.x, r0a.y, lds_f64(a.rows + 16), 0.0}, rowB{r1a.x, r1a.y, lds_f64(a.rows + 48), 0.0};
    unsigned argB[S];
    const int ka = seg_step_margin<S>(V, ea, rowA, a.c0, a.c1, r.acc, a.va, a.chain, a.piece, a.smp, a.obs, a.rec);
    const int kb = seg_step_margin<S>(V, eb, rowB, a.c0, a.c1, argB, a.va, a.chain, a.piece, a.smp, a.obs + 1, a.rec);
#pragma unroll
    for (int j = 0; j < S; j++) {
        r.acc[j] |= argB[j] << 4;
        r.v[j] = V[j];
    }
    r.kinds = (unsigned)ka | (unsigned)kb << 1;
    return r;
}

// one observation of such a piece outside the speculative loop (the ragged ends of a chain; out of line)
template <int S>
struct MStepArgs {
    double v[S], em[S];
    double b0, sf, ot, c0, c1;
    const ViterbiArgs* va;
    int chain, piece, smp, obs;
    int rec;
};
template <int S>
struct MStepRes {
    double v[S];
    unsigned arg[S];
    int kind;
};
template <int S>
__device__ __noinline__ MStepRes<S> tpc_step_margin(const MStepArgs<S> a)
{
    MStepRes<S> r;
    double V[S], em[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        V[j] = a.v[j];
        em[j] = a.em[j];
    }
    const StructRow row{a.b0, a.sf, a.ot, 0.0};
    r.kind = seg_step_margin<S>(V, em, row, a.c0, a.c1, r.arg, a.va, a.chain, a.piece, a.smp, a.obs, a.rec);
#pragma unroll
    for (int j = 0; j < S; j++) r.v[j] = V[j];
    return r;
}

// a work item of the sweep: a whole (chain, 32 samples) line of tiles, or one piece of it (segmented sweep)
struct TpcItem {
    int chain, g32;
    int t_begin, t_seam, t_end;     // tiles of the chain: swept from t_begin, recorded from t_seam, up to t_end
    int piece;
    bool skip;
};
template <bool kSeg>
__device__ __forceinline__ TpcItem tpc_item(const ViterbiArgs& a, int it)
{
    TpcItem wi;
    if (kSeg) {
        wi.piece = a.sched_items[2 * it];
        const int4 d = a.seg_desc[wi.piece];
        wi.chain = d.x;
        wi.g32 = d.y;
        wi.t_seam = d.z;
        wi.t_end = d.w;
        wi.t_begin = d.z > 0 ? d.z - a.seg_warm : 0;
        wi.skip = false;
    } else {
        wi.piece = 0;
        wi.chain = a.sched_items[2 * it];
        wi.g32 = a.sched_items[2 * it + 1];
        wi.t_begin = wi.t_seam = 0;
        wi.t_end = chain_tiles(a.chains[wi.chain]);
        // the repair pass sweeps only the lines the check kernel refused
        wi.skip = a.only_bad && a.seg_flags[seg_off_line(a.n_chains, a.n_samples, wi.chain, wi.g32)] == 0;
    }
    return wi;
}

// kSeg (segmented sweep): no producer warp — every sweep warp issues its own loads kStages half-tiles ahead (the refill
// of a stage follows the warp barrier behind its last read), so that W = 8 puts two sweep warps on every SM
// sub-partition with 255 registers each: a ninth warp would leave one sub-partition's register file to three warps.
// Pieces need throughput, not latency: one sweep warp per sub-partition issues on 30 % of the cycles.
template <int S, int W, bool kSeg>
__global__ void __launch_bounds__((W + (kSeg ? 0 : 1)) * 32, 1)
viterbi_tpc_kernel(const __grid_constant__ ViterbiArgs a, const __grid_constant__ CUtensorMap ll_map)
{
    constexpr int G = 32 / S;                               // chains per back-pointer record (viterbi_common.cuh)
    constexpr int kStages = tpc_stages(S, W);
    static_assert(kStages >= 2, "ring needs two stages");
    constexpr unsigned kStageBytes = tpc_stage_bytes(S);
    constexpr unsigned kEmBytes = tpc_em_bytes(S);
    constexpr int kPairUnroll = kSeg ? 1 : kHalf / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)W * kStages * kStageBytes;
    if (!kSeg && a.only_bad && a.seg_flags[1] == 0) return;    // repair pass of a segmented sweep, nothing refused

    if (threadIdx.x == 0) {
        for (int s = 0; s < W * 2 * kStages; s++) mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (!kSeg && warp == W) {
        // ------------------------------------------------------------------------------------ producer
        if (lane >= W) return;
        const int slot = blockIdx.x * W + lane;
        const uint32_t ring = smem_u32(smem) + (uint32_t)lane * kStages * kStageBytes;
        const uint32_t full = bar0 + (uint32_t)lane * 2 * kStages * 8, empty = full + kStages * 8;
        int st = 0;
        unsigned wrap = 0;
        for (int it = a.sched_begin[slot]; it < a.sched_begin[slot + 1]; it++) {
            const TpcItem wi = tpc_item<kSeg>(a, it);
            if (wi.skip) continue;
            const int chain = wi.chain, g32 = wi.g32;
            const ChainDesc cd = a.chains[chain];
            const int64_t t_first = ((cd.em_off + 1) >> 4) + wi.t_begin;
            const StructRow* __restrict__ rows = a.srows + cd.lt_row0;
            int i0 = (int)((t_first << 4) - cd.em_off);
            int c0 = (int)(t_first << 4);
            const int c1 = g32 * 32 * S;
            for (int t = 2 * wi.t_begin; t < 2 * wi.t_end; t++, i0 += kHalf, c0 += kHalf) {
                if (wrap) mbar_wait_relaxed(empty + 8u * st, (wrap - 1) & 1);
                const int r0 = i0 < 0 ? 0 : i0;
                const int n_rows = i0 + kHalf - r0;         // rows before the chain's first row are never used
                const unsigned row_bytes = n_rows > 0 ? (unsigned)n_rows * (unsigned)sizeof(StructRow) : 0u;
                const uint32_t dst = ring + (uint32_t)st * kStageBytes;
                mbar_expect_tx(full + 8u * st, row_bytes + kEmBytes);
                tma_load_2d(dst, &ll_map, c0, c1, full + 8u * st);
                if (row_bytes)
                    tma_load_1d(dst + kEmBytes + (uint32_t)(r0 - i0) * (uint32_t)sizeof(StructRow), rows + r0, row_bytes, full + 8u * st);
                if (++st == kStages) { st = 0; wrap++; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------- consumer
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * kStages * kStageBytes;
    const uint32_t full = bar0 + (uint32_t)warp * 2 * kStages * 8, empty = full + kStages * 8;
    // this lane's S rows of the emission half-tile (likelihood-column order inside the box); 64-byte rows, 64-byte
    // swizzle: chunk c (16 bytes) of row r sits at chunk c ^ ((r >> 1) & 3), so the 128-bit reads of a quarter warp
    // (8 lanes, 8 distinct r & 7 because S is odd) cover the 32 banks once
    uint32_t em_row[S], em_sw[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        const int r = lane * S + a.perm[j];
        em_row[j] = (uint32_t)r * 64u;
        em_sw[j] = (uint32_t)((r >> 1) & 3) << 4;
    }
    const double c0 = a.c0, c1 = a.c1, tail_other = a.tail_other;
    const double c0m = __dadd_rn(c0, -kSpecMargin);     // the speculative step's acceptance margin

    int st = 0;
    unsigned phase = 0;
    const int slot = blockIdx.x * W + warp;
    // the warp's own feed (kSeg): the half-tile kStages ahead of the one being swept, walking the same work list
    int f_it = a.sched_begin[slot], f_t = 0, f_tend = 0, f_i0 = 0, f_c0 = 0, f_c1 = 0, f_st = 0;
    const int f_end = a.sched_begin[slot + 1];
    const StructRow* __restrict__ f_rows = nullptr;
    auto feed = [&]() {
        while (f_t == f_tend) {
            if (f_it == f_end) return;
            const TpcItem fi = tpc_item<kSeg>(a, f_it++);
            const ChainDesc fd = a.chains[fi.chain];
            const int64_t t_first = ((fd.em_off + 1) >> 4) + fi.t_begin;
            f_rows = a.srows + fd.lt_row0;
            f_i0 = (int)((t_first << 4) - fd.em_off);
            f_c0 = (int)(t_first << 4);
            f_c1 = fi.g32 * 32 * S;
            f_t = 2 * fi.t_begin;
            f_tend = 2 * fi.t_end;
        }
        if (lane == 0) {
            const int r0 = f_i0 < 0 ? 0 : f_i0;
            const int n_rows = f_i0 + kHalf - r0;
            const unsigned row_bytes = n_rows > 0 ? (unsigned)n_rows * (unsigned)sizeof(StructRow) : 0u;
            const uint32_t dst = ring + (uint32_t)f_st * kStageBytes;
            mbar_expect_tx(full + 8u * f_st, row_bytes + kEmBytes);
            tma_load_2d(dst, &ll_map, f_c0, f_c1, full + 8u * f_st);
            if (row_bytes)
                tma_load_1d(dst + kEmBytes + (uint32_t)(r0 - f_i0) * (uint32_t)sizeof(StructRow), f_rows + r0, row_bytes, full + 8u * f_st);
        }
        f_t++, f_i0 += kHalf, f_c0 += kHalf;
        f_st = f_st + 1 == kStages ? 0 : f_st + 1;
    };
    if (kSeg) {
#pragma unroll 1
        for (int q = 0; q < kStages; q++) feed();
    }
    for (int it = a.sched_begin[slot]; it < a.sched_begin[slot + 1]; it++) {
        const TpcItem wi = tpc_item<kSeg>(a, it);
        if (wi.skip) continue;
        const int chain = wi.chain, g32 = wi.g32;
        const ChainDesc cd = a.chains[chain];
        const int nobs = cd.nobs;
        const int64_t t_first = (cd.em_off + 1) >> 4;
        const int n_tiles = chain_tiles(cd);
        const int smp = g32 * 32 + lane;
        const bool live = smp < a.n_samples;                // lanes past the batch sweep zero-filled rows and write nothing
        const int grp = smp / G, gg = smp - grp * G;
        uint2* bp_t = reinterpret_cast<uint2*>(a.bp) + (record_base(a, chain, grp, n_tiles) + wi.t_begin) * kRecU2 + gg * S;
        int i0 = (int)(((t_first + wi.t_begin) << 4) - cd.em_off);
        bool ready = false;
        // a piece that does not start its chain (segmented sweep): V is approximate, every decision is certified
        const bool mseg = kSeg && wi.t_seam > 0;
        unsigned mag_v = 0, mag_e = 0;                      // largest (high word << 1) of V / of the emissions since the seam
        unsigned err_a = 0, err_b = 0, max_a = 0, max_b = 0;   // error multipliers of the relative vector (viterbi_step.h: seg_err_step)

        double V[S];
#pragma unroll
        for (int j = 0; j < S; j++) V[j] = j == 0 ? 0.0 : -HUGE_VAL;         // hmm.cpp:46-52 (and the start of a warm-up)

        for (int t = wi.t_begin; t < wi.t_end; t++, bp_t += kRecU2, i0 += kTile) {
            unsigned lo[S], hi[S];
            const bool rec = !kSeg || t >= wi.t_seam;        // warm-up tiles are swept, not recorded
            if (kSeg && mseg && t == wi.t_seam) {
#pragma unroll
                for (int j = 0; j < S; j++) a.seam_in[((int64_t)wi.piece * S + j) * 32 + lane] = V[j];
                mag_v = mag_e = 0;
                err_a = max_a = 1u;                         // the seam's own deviation e0, carried until the first step that cancels it
                err_b = max_b = 0u;
            }
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                if (!ready) mbar_wait(full + 8u * st, phase);
                const uint32_t stage = ring + (uint32_t)st * kStageBytes;
                const uint32_t rows = stage + kEmBytes;
                int st_n = st + 1;
                unsigned phase_n = phase;
                if (st_n == kStages) { st_n = 0; phase_n ^= 1u; }
                ready = false;
                const int ih = i0 + h * kHalf;              // first observation of this half
                const bool more = h == 0 || t + 1 < wi.t_end;
                unsigned word[S];                           // 8 back-pointers (4 bits each) per destination
                if (ih >= 1 && ih + kHalf - 1 <= cd.n_em) { // half entirely inside the real observations
                    // The speculative step leaves ONE BIT per destination and observation (k = j won, or k = 0); the half's
                    // 8 observations x (S - 1) bits are spread into the record's 4-bit back-pointers at its end (bit -> nibble,
                    // times j).  The exact pair (out of line) overrides its two nibbles per destination.
                    constexpr int kBitWords = (S - 1 + 3) / 4;  // bits of destinations 1..4 in word 0, 5..6 in word 1
                    double2 e[S];
                    uint32_t ea[S];
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        ea[j] = stage + em_row[j] + em_sw[j];
                        e[j] = lds_f64x2(ea[j]);
                    }
                    unsigned pb[kBitWords], ovm[S], ovv[S];
#pragma unroll
                    for (int w = 0; w < kBitWords; w++) pb[w] = 0u;
#pragma unroll
                    for (int j = 0; j < S; j++) ovm[j] = ovv[j] = 0u;
                    // kSeg: one pair per trip — ~230 instructions, inside the 6 KB of a sub-partition's L0 instruction cache; the
                    // four-pair body streamed from L1.5 on every trip and the warps issued at the rate of that fetch (a
                    // no_instruction stall at every 128-byte line, 0.25 instructions per cycle and sub-partition)
#pragma unroll kPairUnroll
                    for (int pp = 0; pp < kHalf / 2; pp++) {
                        double2 en[S];
                        uint32_t ean[S];
                        const uint32_t pn = (uint32_t)((pp + 1) & 3) << 4;   // (the last pair reloads pair 0: harmless)
#pragma unroll
                        for (int j = 0; j < S; j++) {
                            ean[j] = stage + em_row[j] + (pn ^ em_sw[j]);
                            en[j] = lds_f64x2(ean[j]);
                        }
                        const uint32_t ra = rows + (uint32_t)pp * 2u * (uint32_t)sizeof(StructRow);
                        const double2 r0a = lds_f64x2(ra), r1a = lds_f64x2(ra + 32);
                        const double r0o = lds_f64(ra + 16), r1o = lds_f64(ra + 48);
                        if (pp == 2 && more) ready = try_wait_once(full + 8u * st_n, phase_n);   // poll the next stage early
                        unsigned worst = (unsigned)__double2hiint(V[0]) << 1;
                        double emA[S], emB[S], Vn[S];
                        if (kSeg) {
                            unsigned we = 0;
#pragma unroll
                            for (int j = 0; j < S; j++) {
                                we = max(we, max((unsigned)__double2hiint(e[j].x) << 1, (unsigned)__double2hiint(e[j].y) << 1));
                                mag_v = max(mag_v, (unsigned)__double2hiint(V[j]) << 1);
                            }
                            mag_e = max(mag_e, we);
                            worst = max(worst, we);
                        }
#pragma unroll
                        for (int j = 0; j < S; j++) {
                            if (!kSeg) worst = max(worst, max((unsigned)__double2hiint(e[j].x) << 1, (unsigned)__double2hiint(e[j].y) << 1));
                            emA[j] = e[j].x;
                            emB[j] = e[j].y;
                            Vn[j] = V[j];
                        }
                        const StructRow rowA{r0a.x, r0a.y, r0o, 0.0}, rowB{r1a.x, r1a.y, r1o, 0.0};
                        bool ok = worst < kSpecBigHi2;
                        unsigned bitsA, bitsB;
                        if (kSeg) {
                            // the lead of every decision rides along; a pair with a lead below 2^-14 goes the slow way (for a
                            // piece that starts its chain the slow way is simply the exact pair)
                            unsigned min_hi = 0x7FFFFFFFu;
                            bitsA = viterbi_step_spec_m<S>(Vn, emA, c0, c1, c0m, rowA, ok, min_hi);
                            bitsB = viterbi_step_spec_m<S>(Vn, emB, c0, c1, c0m, rowB, ok, min_hi);
                            ok = ok && min_hi >= kSegTauHi;
                            if (mseg && ok) {
                                // destination 0 keeps k = 0 here; no bit set = every destination was won by k = 0: the deviations cancel
                                err_a = bitsA ? err_a : 0u;
                                err_b = bitsA ? err_b + 1u : 1u;
                                max_b = max(max_b, err_b);
                                err_a = bitsB ? err_a : 0u;
                                err_b = bitsB ? err_b + 1u : 1u;
                                max_b = max(max_b, err_b);
                            }
                        } else {
                            bitsA = viterbi_step_spec<S>(Vn, emA, c0, c1, c0m, rowA, ok);
                            bitsB = viterbi_step_spec<S>(Vn, emB, c0, c1, c0m, rowB, ok);
                        }
#pragma unroll
                        for (int w = 0; w < kBitWords; w++)
                            pb[w] |= ((bitsA >> (4 * w)) & 0xFu) << (8 * pp) | ((bitsB >> (4 * w)) & 0xFu) << (8 * pp + 4);
                        if (!ok && (!(kSeg && mseg) || live)) {
                            // (dead lanes of a certified piece sweep zero-filled rows — ties everywhere — and stay out of the slow path)
                            PairArgs<S> pa;
#pragma unroll
                            for (int j = 0; j < S; j++) {
                                pa.v[j] = V[j];
                                pa.em[j] = ea[j];
                            }
                            pa.rows = ra;
                            pa.c0 = c0;
                            pa.c1 = c1;
                            pa.va = &a;
                            pa.chain = chain, pa.piece = wi.piece, pa.smp = smp, pa.obs = ih + 2 * pp, pa.rec = rec ? 1 : 0;
                            PairRes<S> pr;
                            if (kSeg && mseg) {
                                pr = tpc_pair_margin<S>(pa);
                                seg_err_step(err_a, err_b, (int)(pr.kinds & 1u));
                                max_b = max(max_b, err_b);
                                seg_err_step(err_a, err_b, (int)(pr.kinds >> 1));
                                max_b = max(max_b, err_b);
                            } else pr = tpc_pair_exact<S>(pa);
#pragma unroll
                            for (int j = 0; j < S; j++) {
                                Vn[j] = pr.v[j];
                                ovm[j] |= 0xFFu << (8 * pp);
                                ovv[j] |= pr.acc[j] << (8 * pp);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < S; j++) {
                            V[j] = Vn[j];
                            e[j] = en[j];
                            ea[j] = ean[j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        // destination j's bit of every observation -> its nibble, times j; destination 0: all zero
                        const unsigned spread = j == 0 ? 0u : ((pb[(j - 1) / 4] >> ((j - 1) & 3)) & 0x11111111u) * (unsigned)j;
                        word[j] = (spread & ~ovm[j]) | ovv[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < S; j++) word[j] = 0u;
#pragma unroll 1
                    for (int q = 0; q < kHalf; q++) {
                        const int i = ih + q;
                        unsigned arg[S];
#pragma unroll
                        for (int j = 0; j < S; j++) arg[j] = (unsigned)j;    // observations outside the chain: identity step
                        if (i >= 1 && i < nobs) {           // warp-uniform
                            double em[S];
                            const uint32_t qa = ((uint32_t)(q >> 1) << 4), qb = (uint32_t)(q & 1) << 3;
#pragma unroll
                            for (int j = 0; j < S; j++) {
                                em[j] = lds_f64(stage + em_row[j] + (qa ^ em_sw[j]) + qb);
                                if (i > cd.n_em) em[j] = j == 0 ? 0.0 : tail_other;
                            }
                            const uint32_t ra = rows + (uint32_t)q * (uint32_t)sizeof(StructRow);
                            const double2 rab = lds_f64x2(ra);
                            const StructRow row{rab.x, rab.y, lds_f64(ra + 16), 0.0};
                            if (kSeg && mseg) {
                                if (live) {
                                    MStepArgs<S> ma;
#pragma unroll
                                    for (int j = 0; j < S; j++) {
                                        ma.v[j] = V[j];
                                        ma.em[j] = em[j];
                                        mag_v = max(mag_v, (unsigned)__double2hiint(V[j]) << 1);
                                        mag_e = max(mag_e, (unsigned)__double2hiint(em[j]) << 1);
                                    }
                                    ma.b0 = row.b0, ma.sf = row.sf, ma.ot = row.ot, ma.c0 = c0, ma.c1 = c1;
                                    ma.va = &a;
                                    ma.chain = chain, ma.piece = wi.piece, ma.smp = smp, ma.obs = i, ma.rec = rec ? 1 : 0;
                                    const MStepRes<S> rm = tpc_step_margin<S>(ma);
                                    seg_err_step(err_a, err_b, rm.kind);
                                    max_b = max(max_b, err_b);
#pragma unroll
                                    for (int j = 0; j < S; j++) {
                                        V[j] = rm.v[j];
                                        arg[j] = rm.arg[j];
                                    }
                                }
                            } else viterbi_step_struct<S>(V, em, c0, c1, row, arg);
                        }
#pragma unroll
                        for (int j = 0; j < S; j++) word[j] |= arg[j] << (4 * q);
                    }
                }
#pragma unroll
                for (int j = 0; j < S; j++) {
                    if (h == 0) lo[j] = word[j];
                    else hi[j] = word[j];
                }
                __syncwarp();                               // every lane is done with the stage: hand it back
                if (kSeg) feed();                           // (... to this warp's own feed: the stage takes the half-tile kStages ahead)
                else if (lane == 0) mbar_arrive(empty + 8u * st);
                st = st_n;
                phase = phase_n;
            }
            if (live && rec) {
#pragma unroll
                for (int j = 0; j < S; j++) bp_t[j] = make_uint2(lo[j], hi[j]);
            }
        }
        if (kSeg) {
            // what the check kernel needs to certify the NEXT piece of this line (and this one): V after the last
            // observation, and how large the values of this piece were
#pragma unroll
            for (int j = 0; j < S; j++) {
                a.seam_out[((int64_t)wi.piece * S + j) * 32 + lane] = V[j];
                mag_v = max(mag_v, (unsigned)__double2hiint(V[j]) << 1);
            }
            unsigned* pe = a.seam_mag + (int64_t)wi.piece * (kSeamWords * 32) + lane;
            pe[0] = mag_v, pe[32] = mag_e, pe[64] = max_a, pe[96] = max_b, pe[128] = err_a, pe[160] = err_b;
            // this piece's share of the line's bound of |C| (viterbi_seam.h)
            double* cabs = reinterpret_cast<double*>(a.seg_flags + seg_off_cabs(a.n_chains, a.n_samples));
            // (the piece that starts its chain is exact and enters the sum through its last V[0] alone: its V[k > 0] start at -Inf)
            atomicAdd(cabs + ((int64_t)chain * seg_n_g32(a.n_samples) + g32) * 32 + lane,
                      piece_cabs_share(mseg ? mag_v : (unsigned)__double2hiint(V[0]) << 1, (wi.t_end - wi.t_seam) * kTile));
        }
    }
}

// ---- check kernel of the segmented sweep (after expand: it reads the path) --------------------------------------------
// Part 1, one thread per (piece, sample): the seam in front of the piece (viterbi_seam.h: seam_check) — the chain is
// refused when the seam does not close or the certified deviation outgrows kSegEpsMax.  Part 2, one thread per listed
// decision (lead below kSegTau): certified after all when its lead exceeds the deviation its piece actually reached
// (typically 1e-7 against leads spread over 0 .. 6e-5); otherwise the decision (observation i, destination j) is read by
// the traceback only if the path is in state j at observation i — then the chain is refused too.  Refused chains are swept
// again by the repair pass.
template <int S>
__global__ void __launch_bounds__(128)
viterbi_seg_check_kernel(ViterbiArgs a, int n_pieces)
{
    const int n_g32 = seg_n_g32(a.n_samples);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, n_thr = (int64_t)gridDim.x * blockDim.x;
    const double* cabs = reinterpret_cast<const double*>(a.seg_flags + seg_off_cabs(a.n_chains, a.n_samples));
    // the seam in front of piece pc for lane `lane`: 0 / the reason it is refused; *certified: the lead above which the
    // piece's decisions are certified
    auto seam = [&](int pc, int lane, const int4& d, double* certified) -> int {
        double x_in[S], x_prev[S];
#pragma unroll
        for (int j = 0; j < S; j++) {
            x_in[j] = a.seam_in[((int64_t)pc * S + j) * 32 + lane];
            x_prev[j] = a.seam_out[((int64_t)(pc - 1) * S + j) * 32 + lane];
        }
        const unsigned* w = a.seam_mag + (int64_t)pc * (kSeamWords * 32) + lane;
        const PieceErr pe{w[0], w[32], w[64], w[96], w[128], w[160]};
        const unsigned* v = w - kSeamWords * 32;
        const PieceErr prev{v[0], v[32], v[64], v[96], v[128], v[160]};
        const bool prev_exact = a.seg_desc[pc - 1].z == 0;          // piece pc - 1 is the one before it on its line
        return seam_check<S>(x_in, x_prev, pe, prev_exact ? nullptr : &prev, cabs[((int64_t)d.x * n_g32 + d.y) * 32 + lane], certified);
    };
    for (int64_t q = tid; q < (int64_t)n_pieces * 32; q += n_thr) {
        const int pc = (int)(q >> 5), lane = (int)(q & 31);
        const int4 d = a.seg_desc[pc];
        const int smp = d.y * 32 + lane;
        if (smp >= a.n_samples) continue;
        int bad = a.seg_force_repair ? kBadForced : 0;
        if (d.z > 0) bad |= seam(pc, lane, d, nullptr);     // a piece behind a seam
        if (bad) seg_mark_bad(a.seg_flags, a.n_chains, a.n_samples, d.x, smp, bad);
    }
    const int n_close = min(a.seg_flags[0], a.seg_close_cap);
    for (int64_t q = tid; q < n_close; q += n_thr) {
        const int4 e = a.seg_close[q];                      // (piece, sample, observation, destination | lead)
        const int4 d = a.seg_desc[e.x];
        const int dest = e.w & 7;
        const double lead = (double)__uint_as_float((unsigned)e.w & ~7u);
        double certified = 0.0;
        if (seam(e.x, e.y & 31, d, &certified)) continue;   // a refused seam: part 1 has refused the chain already
        if (lead > certified) continue;                     // listed, yet far enough above what the piece's deviation reached
        const ChainDesc cd = a.chains[d.x];
        bool on = true;                                     // an observation whose state is not on record: assume the worst
        if (e.z == cd.nobs - 1) on = dest == 0;             // the chain ends in state 0 (hmm.cpp:96)
        else if (e.z >= cd.out_first && e.z <= cd.out_last) on = (int)a.path[e.y * a.path_stride + cd.out_off + e.z] == dest;
        if (on) seg_mark_bad(a.seg_flags, a.n_chains, a.n_samples, d.x, e.y, kBadOnPath);
    }
}

int launch_viterbi_seg_check(const ViterbiArgs& a, int n_pieces, cudaStream_t st)
{
    const int64_t n = (int64_t)n_pieces * 32;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, 2368));
    switch (a.n_states) {
        case 3: viterbi_seg_check_kernel<3><<<blocks, 128, 0, st>>>(a, n_pieces); break;
        case 5: viterbi_seg_check_kernel<5><<<blocks, 128, 0, st>>>(a, n_pieces); break;
        case 7: viterbi_seg_check_kernel<7><<<blocks, 128, 0, st>>>(a, n_pieces); break;
        default: return 0;
    }
    return 1;
}

size_t viterbi_tpc_smem_bytes(int S, int W) { return (size_t)W * tpc_stages(S, W) * (tpc_stage_bytes(S) + 16); }
int viterbi_tpc_max_warps(int S) { return 4; }
// sweep warps per CTA of the segmented sweep: two per SM sub-partition where two ring stages per warp fit the shared memory
int viterbi_seg_warps(int S) { return tpc_stages(S, 8) >= 2 ? 8 : tpc_stages(S, 6) >= 2 ? 6 : 4; }

template <int S, int W, bool kSeg>
static void launch_tpc(const ViterbiArgs& a, cudaStream_t st)
{
    if constexpr (tpc_stages(S, W) >= 2) {
        const size_t smem = viterbi_tpc_smem_bytes(S, W);
        static PerDevice configured;
        if (configured.raise(smem)) cudaFuncSetAttribute(viterbi_tpc_kernel<S, W, kSeg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        viterbi_tpc_kernel<S, W, kSeg><<<a.n_slots / W, (W + (kSeg ? 0 : 1)) * 32, smem, st>>>(a, *reinterpret_cast<const CUtensorMap*>(a.ll_map_tpc));
    }
}

template <int S>
static void launch_tpc_w(const ViterbiArgs& a, cudaStream_t st)
{
    if (a.seg) {                                            // pieces are dealt for 4, 6 or 8 self-feeding sweep warps per CTA
        if (a.warps_per_cta == 8) launch_tpc<S, 8, true>(a, st);
        else if (a.warps_per_cta == 6) launch_tpc<S, 6, true>(a, st);
        else launch_tpc<S, 4, true>(a, st);
        return;
    }
    switch (a.warps_per_cta) {
        case 1: launch_tpc<S, 1, false>(a, st); break;
        case 2: launch_tpc<S, 2, false>(a, st); break;
        case 3: launch_tpc<S, 3, false>(a, st); break;
        default: launch_tpc<S, 4, false>(a, st); break;
    }
}

// the structured sweep of the chains in a.chain_list (schedule: 32-sample groups); returns the number of launches
int launch_viterbi_tpc_sweep(const ViterbiArgs& a, cudaStream_t st)
{
    switch (a.n_states) {
        case 3: launch_tpc_w<3>(a, st); break;
        case 5: launch_tpc_w<5>(a, st); break;
        case 7: launch_tpc_w<7>(a, st); break;
        default: return 0;
    }
    return 1;
}

}  // namespace edb
