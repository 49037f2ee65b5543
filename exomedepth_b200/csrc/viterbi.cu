// viterbi.cu — HMM Viterbi sweep, traceback and segment summary (sm_100a).
//
// Replaces C_hmm of the reference (src/hmm.cpp:18-167).  The recurrence is the reference's, term for
// term: candidates (emis + V[k]) + log(t_k) in that association (hmm.cpp:79), strict '>' so the
// lowest source state wins ties and NaN candidates are skipped (hmm.cpp:81), from = 0 when the
// emission is -Inf (hmm.cpp:87), forced end in state 0 (hmm.cpp:96), and the segment scan with its
// stale `start` on a direct CNV->CNV change (hmm.cpp:104-126).  The device performs only IEEE
// add / compare on FP64, which are exact, so the backtrace is bit-identical to the reference given
// the same emissions and the same log-transition table (built on the host with the host libm, see
// host_tables.cpp) — no FMA contraction, no (max,+) re-association.
//
// Five kernels (DESIGN.md "Viterbi"):
//   sweep    one lane per (chain, destination state); a warp carries G = 32/S chains of the SAME
//            chromosome from G consecutive samples.  The sweep is a chain of dependent FP64 operations
//            (~20,000 steps for chromosome 1), so this kernel is about the latency of ONE step
//            (tools/ubench/step.cu measures the variants on a B200).  Warp roles: 4 or 8 sweep warps and one
//            producer warp per CTA; the producer issues, per 16-observation tile and sweep warp, one 2-D TMA
//            load of the emission tile (128-byte swizzle) and one bulk copy of the tile's log-transition rows
//            into the warp's ring stage (full / empty mbarriers).  Inside a tile the S lanes of a chain exchange
//            V[i-1][k] through shared memory (STS.64, warp barrier, 128-bit loads), the additions are pinned
//            behind the last exchange load, transition rows and emissions are read one step ahead, the maximum
//            is a left-keeping tournament whose predicates also select the winning source state (the
//            back-pointer), and back-pointers leave as 64 bits per lane and tile.
//            Work items (chromosome x group of samples) are placed on the SM sub-partitions by the
//            host (longest-processing-time first), so the longest chains run alone on theirs.
//            This file is compiled with ptxas -O1: at the default level ptxas interleaves the exchange
//            with its consumers (2.32 instead of 2.06 ms per sweep when measured).
//   tilemap  composes the 16 back-pointer steps of every tile into a map end state -> state before
//            the tile (fully parallel);
//   trace    one warp per chain walks one map per tile instead of one back-pointer per observation
//            (hmm.cpp:95-100);
//   expand   one warp per chain fills in the per-observation states, the path bytes and the
//            reference's call table, 32 tiles at a time;
//   compact  concatenates the per-chromosome call tables per sample.
#include <algorithm>
#include <vector>

#include <cuda.h>

#include "host_tables.h"
#include "kernels.cuh"
#include "tma_ptx.cuh"
#include "viterbi_common.cuh"

namespace edb {

// Sweep warps per CTA (one more warp issues the TMA loads): 4 = one per SM sub-partition, 8 = two.  A sweep warp alone
// on its sub-partition AND in a CTA of 4 steps in ~135 cycles; with 8 per CTA the warps contend for the SM's
// shared-memory / shuffle pipe and each step takes ~180 (tools/ubench/step.cu) — 8 only pays when the batch has more
// work than 4 warps per SM can turn over within the longest chromosome's chain (viterbi_pick_warps).
constexpr int kEmBytes = 4096;     // emission tile: up to 32 rows x 128 bytes, 1024-byte aligned for the 128-byte swizzle

// lt_jstride / lt_pitch (layout of a log-transition row): host_tables.h, shared with the host-side table builder
// one ring stage of a sweep warp: [emission tile 4 KB][transition rows of the tile's 16 observations], 1 KB granular
__host__ __device__ constexpr int stage_bytes(int S) { return (kEmBytes + kTile * lt_pitch(S) * 8 + 1023) / 1024 * 1024; }
// ring depth: what fits in ~220 KB of shared memory per CTA, at most 6 stages
__host__ __device__ constexpr int ring_stages(int S, int W)
{
    const int fit = (220 * 1024) / (W * stage_bytes(S));
    return fit > 6 ? 6 : fit;
}

// ---- the recurrence -----------------------------------------------------------------------------------
// cand_k = (em + V[k]) + lt[k]  (hmm.cpp:79), winner = FIRST maximum (strict '>' of hmm.cpp:81).
// The maximum is taken with an order-preserving tournament on the values: a node keeps its left entry
// unless the right one is strictly greater, so the surviving VALUE is the one the reference's sequential
// scan keeps, and the winning source state is the lowest k whose candidate equals it.
// NaN never reaches the tournament: NaN transition terms are stored as -Inf in the device copy of the
// table (a NaN candidate and a -Inf candidate are both "never selected", hmm.cpp:81) and a NaN emission
// is replaced by -Inf (all candidates lose, from stays -1, V stays -Inf, as in the reference).
template <int S>
struct Cand {
    int arg;                        // source state of the step's maximum (first maximum, like the reference's scan)
};

// what one step reads from shared memory: this lane's transition row and its emission.  Loaded one step AHEAD
// of its use (the warp issues in order: a load placed next to its consumer stalls the whole dependent chain).
template <int S>
struct StepIn {
    double lt[S + 1];
    double em;
};
template <int S>
__device__ __forceinline__ void load_step(StepIn<S>& in, uint32_t lt_qj, uint32_t em_addr)
{
#pragma unroll
    for (int k = 0; k + 1 < S; k += 2) {
        const double2 x = lds_f64x2(lt_qj + 8u * k);
        in.lt[k] = x.x;
        in.lt[k + 1] = x.y;
    }
    if (S & 1) in.lt[S - 1] = lds_f64(lt_qj + 8u * (S - 1));
    in.em = lds_f64(em_addr);
}

// exchange V between the chain's S lanes, form the candidates and reduce them to the new V.
// The exchange goes through shared memory (one STS.64, a warp barrier, then 128-bit loads of the chain's S values)
// rather than through 2*S shuffles: the latency is the same (tools/ubench/step.cu: 126 cycles per step either way)
// but the warp barrier pins the order — ptxas is free to sink individual shuffles next to their consumers, which
// exposes one shuffle latency per source state on the dependent chain (measured: 178 instead of 130 cycles).
// `xch` = this chain's slot for this step (double-buffered by step parity), `xch_own` = this lane's entry in it.
template <int S>
__device__ __forceinline__ void sweep_step(const StepIn<S>& in, uint32_t xch, uint32_t xch_own, double& V, Cand<S>& cd)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(xch_own), "d"(V) : "memory");
    __syncwarp();
    double v[S + 1];
#pragma unroll
    for (int k = 0; k + 1 < S; k += 2) {
        const double2 x = lds_f64x2_fresh(xch + 8u * k);
        v[k] = x.x;
        v[k + 1] = x.y;
    }
    if (S & 1) v[S - 1] = lds_f64_fresh(xch + 8u * (S - 1));
    double em_s = in.em != in.em ? -HUGE_VAL : in.em;               // depends on the emission only: off the chain through V
    // Order pin: ptxas likes to place the first addition right behind the first of the loads above, which stalls the
    // warp there with the other loads not yet issued (one extra shared-memory latency per step).  Tying em_s to the
    // LAST load — through a predicate that is never true: V is never a NaN, let alone this one — makes every
    // addition wait until all loads are in flight.
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0x7ff1d00d;\n\t@p mov.f64 %0, 0d0000000000000000;\n\t}" : "+d"(em_s) : "r"(__double2hiint(v[S - 1])));
    double m[S];
    int id[S];
#pragma unroll
    for (int k = 0; k < S; k++) {
        m[k] = __dadd_rn(__dadd_rn(em_s, v[k]), in.lt[k]);
        id[k] = k;
    }
    // pairs first; the last three survivors are settled by independent compares.  A node keeps its left entry unless
    // the right one is strictly greater, so the surviving index is the reference's first maximum; the index selects
    // ride on the predicates the value selects need anyway and feed nothing on the dependent chain.
    int n = S;
#pragma unroll
    for (; n > 3; n = (n + 1) / 2) {
#pragma unroll
        for (int p = 0; p + 1 < n; p += 2) {
            const bool right = m[p + 1] > m[p];
            m[p / 2] = right ? m[p + 1] : m[p];
            id[p / 2] = right ? id[p + 1] : id[p];
        }
        if (n & 1) { m[n / 2] = m[n - 1]; id[n / 2] = id[n - 1]; }
    }
    if (n == 3) {
        const bool p = m[1] > m[0], q2 = m[2] > m[0], r2 = m[2] > m[1];
        const double t = p ? m[1] : m[0];
        const int ti = p ? id[1] : id[0];
        V = (q2 && r2) ? m[2] : t;
        cd.arg = (q2 && r2) ? id[2] : ti;
    } else if (n == 2) {
        const bool p = m[1] > m[0];
        V = p ? m[1] : m[0];
        cd.arg = p ? id[1] : id[0];
    } else {
        V = m[0];
        cd.arg = id[0];
    }
}

// back-pointer of the step whose winning source is in `cd` and whose maximum is V
template <int S>
__device__ __forceinline__ unsigned sweep_arg(const Cand<S>& cd, double V, double em)
{
    unsigned arg = (unsigned)cd.arg;
    if (!(V > -HUGE_VAL)) arg = 7u;                         // 7 encodes "from = -1" (hmm.cpp:60)
    if (em == -HUGE_VAL) arg = 0u;                          // hmm.cpp:87
    return arg;
}

// =========================================================================================== sweep
// Warp roles: warps 0..7 sweep (consumers), warp 8 feeds them (producer): lane w of the producer walks the work
// list of sweep warp w one tile ahead of it and issues, per tile, one 2-D TMA load of the emission tile and one
// bulk copy of the tile's transition rows into the warp's ring stage, both completing on the stage's `full`
// mbarrier; the sweep warp releases the stage through its `empty` mbarrier.  The sweep warps therefore carry no
// address arithmetic, no expect_tx and no TMA issue between two tiles of their dependent chain.
template <int S, int kWarpsPerCta>
__global__ void __launch_bounds__((kWarpsPerCta + 1) * 32, 1)
viterbi_sweep_kernel(ViterbiArgs a, const __grid_constant__ CUtensorMap ll_map)
{
    constexpr int G = 32 / S;
    constexpr int LTP = lt_pitch(S);
    constexpr int LTJ = lt_jstride(S);
    constexpr int kStages = ring_stages(S, kWarpsPerCta);
    constexpr unsigned kStageBytes = stage_bytes(S);
    constexpr unsigned kEmBox = G * S * kTile * 8;          // bytes one emission tile delivers
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [stages: warp x stage x kStageBytes][mbarriers: warp x (full[kStages], empty[kStages])][V exchange: warp x 2 x G x LTJ doubles]
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)kWarpsPerCta * kStages * kStageBytes;
    const uint32_t xch0 = bar0 + (uint32_t)kWarpsPerCta * 2 * kStages * 8;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWarpsPerCta * 2 * kStages; s++) mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                        // the only CTA-wide barrier: every warp is its own pipeline below

    if (warp == kWarpsPerCta) {
        // ------------------------------------------------------------------------------------ producer
        if (lane >= kWarpsPerCta) return;
        const int slot = blockIdx.x * kWarpsPerCta + lane;
        const uint32_t ring = smem_u32(smem) + (uint32_t)lane * kStages * kStageBytes;
        const uint32_t full = bar0 + (uint32_t)lane * 2 * kStages * 8, empty = full + kStages * 8;
        int st = 0;
        unsigned wrap = 0;                                  // completed passes over the ring
        for (int it = a.sched_begin[slot]; it < a.sched_begin[slot + 1]; it++) {
            const int chain = a.sched_items[2 * it], grp = a.sched_items[2 * it + 1];
            const ChainDesc cd = a.chains[chain];
            const int64_t t_first = (cd.em_off + 1) >> 4;
            const int n_tiles = chain_tiles(cd);
            const double* __restrict__ lt_base = a.lt + cd.lt_row0 * LTP;
            int i0 = (int)((t_first << 4) - cd.em_off);     // first observation of the tile (may be < 1)
            int c0 = (int)(t_first << 4);
            const int c1 = grp * G * S;
            for (int t = 0; t < n_tiles; t++, i0 += kTile, c0 += kTile) {
                if (wrap) mbar_wait_relaxed(empty + 8u * st, (wrap - 1) & 1);
                const int r0 = i0 < 0 ? 0 : i0;             // rows before the chain's first row are never used
                const unsigned lt_bytes = (unsigned)(i0 + kTile - r0) * LTP * 8;
                const uint32_t dst = ring + (uint32_t)st * kStageBytes;
                mbar_expect_tx(full + 8u * st, lt_bytes + kEmBox);
                // emission tile: one 2-D box (16 bins x the G*S rows of the warp's chains; rows past the batch and
                // columns past the matrix are zero-filled by the TMA unit)
                tma_load_2d(dst, &ll_map, c0, c1, full + 8u * st);
                tma_load_1d(dst + kEmBytes + (uint32_t)(r0 - i0) * LTP * 8, lt_base + (int64_t)r0 * LTP, lt_bytes, full + 8u * st);
                if (++st == kStages) { st = 0; wrap++; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------- consumer
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * kStages * kStageBytes;
    const uint32_t full = bar0 + (uint32_t)warp * 2 * kStages * 8, empty = full + kStages * 8;
    int g = lane / S;
    const int j = lane - g * S;
    if (g >= G) g = G - 1;                                  // spare lanes shadow lanes of the last chain
    // V exchange slots of this lane's chain (two, alternating by step) and the lane's own entry in them; the spare
    // lanes write to a slot of their own that nobody reads
    constexpr uint32_t kXchBuf = (G + 1) * LTJ * 8;
    const uint32_t xch = xch0 + (uint32_t)warp * 2 * kXchBuf + (uint32_t)g * LTJ * 8;
    const uint32_t xch_own = (lane < G * S ? xch : xch0 + (uint32_t)warp * 2 * kXchBuf + (uint32_t)G * LTJ * 8) + 8u * j;
    const double tail = j == 0 ? 0.0 : a.tail_other;
    // this lane's row of the emission tile: the box holds the G*S rows of the warp's chains in likelihood-column
    // order; chunk c (16 bytes) of row r sits at chunk c ^ (r & 7) (128-byte swizzle): conflict-free reads
    const int em_r = g * S + a.perm[j];
    const uint32_t em_lane = (uint32_t)em_r * 128u + ((uint32_t)(em_r & 7) << 4);
    const uint32_t lt_lane = kEmBytes + (uint32_t)(j * LTJ) * 8;

    int st = 0;
    unsigned phase = 0;
    const int slot = blockIdx.x * kWarpsPerCta + warp;
    for (int it = a.sched_begin[slot]; it < a.sched_begin[slot + 1]; it++) {
        const int chain = a.sched_items[2 * it], grp = a.sched_items[2 * it + 1];
        const ChainDesc cd = a.chains[chain];
        const int nobs = cd.nobs;
        // tiles follow the 128-byte lines of the emission rows: tile t covers observations i with
        // (em_off + i) / 16 == t_first + t
        const int64_t t_first = (cd.em_off + 1) >> 4;
        const int n_tiles = chain_tiles(cd);
        uint2* bp_t = reinterpret_cast<uint2*>(a.bp) + record_base(a, chain, grp, n_tiles) * kRecU2 + lane;
        int i0 = (int)((t_first << 4) - cd.em_off);         // first observation of the tile (may be < 1)
        bool ready = false;

        double V = j == 0 ? 0.0 : -HUGE_VAL;                // hmm.cpp:46-52

        for (int t = 0; t < n_tiles; t++, bp_t += kRecU2, i0 += kTile) {
            if (!ready) mbar_wait(full + 8u * st, phase);
            const uint32_t stage = ring + (uint32_t)st * kStageBytes;
            const uint32_t ltt = stage + lt_lane;
            const uint32_t emt = stage + em_lane;           // element q: the 128-byte swizzle is an XOR on address bits 4-6
            int st_n = st + 1;
            unsigned phase_n = phase;
            if (st_n == kStages) { st_n = 0; phase_n ^= 1u; }
            ready = false;
            unsigned lo = 0, hi = 0;                        // 16 back-pointers of this lane, 4 bits each
            if (i0 >= 1 && i0 + kTile - 1 <= cd.n_em) {     // tile entirely inside the real observations
                Cand<S> cnd;
                double Vq = V, em_prev = 0.0;
                StepIn<S> cur;
                load_step<S>(cur, ltt, emt);
#pragma unroll
                for (int q = 0; q < kTile; q++) {
                    StepIn<S> nxi;
                    if (q + 1 < kTile) load_step<S>(nxi, ltt + (q + 1) * LTP * 8, emt ^ (uint32_t)((q + 1) << 3));
                    if (q == kTile / 2 && t + 1 < n_tiles) ready = try_wait_once(full + 8u * st_n, phase_n);   // poll the next tile early
                    Cand<S> nxt;
                    double Vn = Vq;
                    sweep_step<S>(cur, xch + (q & 1) * kXchBuf, xch_own + (q & 1) * kXchBuf, Vn, nxt);
                    if (q > 0) {                            // the previous step's back-pointer, in the shadow of this step's exchange
                        const unsigned arg = sweep_arg<S>(cnd, Vq, em_prev);
                        if (q - 1 < 8) lo |= arg << (4 * (q - 1));
                        else hi |= arg << (4 * (q - 9));
                    }
                    cnd = nxt;
                    Vq = Vn;
                    em_prev = cur.em;
                    if (q + 1 < kTile) cur = nxi;
                }
                hi |= sweep_arg<S>(cnd, Vq, em_prev) << 28;
                V = Vq;
            } else {
#pragma unroll 1
                for (int q = 0; q < kTile; q++) {
                    const int i = i0 + q;
                    unsigned arg = (unsigned)j;             // observations outside the chain: identity step
                    if (i >= 1 && i < nobs) {               // warp-uniform
                        StepIn<S> in;
                        load_step<S>(in, ltt + q * LTP * 8, emt ^ (uint32_t)(q << 3));
                        if (i > cd.n_em) in.em = tail;
                        Cand<S> cnd;
                        sweep_step<S>(in, xch + (q & 1) * kXchBuf, xch_own + (q & 1) * kXchBuf, V, cnd);
                        arg = sweep_arg<S>(cnd, V, in.em);
                    }
                    if (q < 8) lo |= arg << (4 * q);
                    else hi |= arg << (4 * (q - 8));
                }
            }
            *bp_t = make_uint2(lo, hi);
            __syncwarp();                                   // every lane is done with the stage: hand it back
            if (lane == 0) mbar_arrive(empty + 8u * st);
            st = st_n;
            phase = phase_n;
        }
    }
}

// =========================================================================================== tilemap
// One thread per (record, chain of the record).  cur holds, for every end state e (nibble e; nibble 7 = the
// pseudo-state "from = -1"), the state reached walking back from the tile's last observation.  As soon as all
// tracked nibbles agree (the usual case: every state's best predecessor is "normal" somewhere in the tile) one
// walk serves all of them.
template <int S>
__global__ void __launch_bounds__(256)
viterbi_tilemap_kernel(ViterbiArgs a)
{
    constexpr int G = 32 / S;
    constexpr unsigned kTracked = (S >= 7 ? 0x0FFFFFFFu : ((1u << (4 * S)) - 1u)) | 0xF0000000u;
    constexpr unsigned kOnes = kTracked & 0x11111111u;
    // the records of this thread: all chains as they lie (flat), or the launch's chain blockIdx.y (grid-stride: the repair
    // pass of a segmented sweep launches a few blocks per chain, which normally find nothing to do)
    int64_t first, count, step = (int64_t)gridDim.x * blockDim.x, base = 0;
    if (a.flat_records > 0) count = a.flat_records * G;
    else {
        if (a.only_bad && a.seg_flags[1] == 0) return;          // repair pass of the segmented sweep, nothing refused
        const int chain = a.chain_list ? a.chain_list[blockIdx.y] : (int)blockIdx.y;
        if (a.only_bad && a.seg_flags[seg_off_chain(chain)] == 0) return;      // ... refused chains only
        count = (int64_t)chain_tiles(a.chains[chain]) * a.groups * G;
        base = (int64_t)a.bp_tile_base[chain] * a.groups;
    }
    first = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t idx = first; idx < count; idx += step) {
    const int64_t r = base + idx / G;
    const int gg = (int)(idx % G);
    const uint2* __restrict__ rec = reinterpret_cast<const uint2*>(a.bp) + r * kRecU2;
    uint2 w[S];
#pragma unroll
    for (int s = 0; s < S; s++) w[s] = rec[gg * S + s];
    unsigned cur = 0x76543210u & kTracked;
    int q = kTile - 1;
    for (; q >= 0 && cur != (cur & 0xFu) * kOnes; q--) {
        unsigned m = 0;                                     // step map: nibble s = predecessor of state s; nibble 7 = 0 (pinned like oracle.c)
#pragma unroll
        for (int s = 0; s < S; s++) m |= bp_nibble(w[s], q) << (4 * s);
        unsigned nc = 0;
#pragma unroll
        for (int e = 0; e < 8; e++)
            if ((kTracked >> (4 * e)) & 1u) nc |= ((m >> (4 * ((cur >> (4 * e)) & 0xFu))) & 0xFu) << (4 * e);
        cur = nc;
    }
    if (q >= 0) {                                           // all end states share one walk from here
        unsigned st1 = cur & 0xFu;
        // the usual case away from CNV regions: the walk has reached state 0 and state 0's own back-pointers are 0 for the
        // rest of the tile — it stays there (the 15 remaining steps were 3/4 of this kernel's instructions)
        if (st1 == 0u) {
            const unsigned long long w0 = ((unsigned long long)w[0].y << 32) | w[0].x;
            if ((w0 & (q >= 15 ? ~0ull : (1ull << (4 * (q + 1))) - 1ull)) == 0ull) q = -1;
        }
        for (; q >= 0; q--) {
            uint2 ws = w[0];
#pragma unroll
            for (int s = 1; s < S; s++) ws = st1 == (unsigned)s ? w[s] : ws;
            st1 = st1 == 7u ? 0u : bp_nibble(ws, q);
        }
        cur = st1 * kOnes;
    }
    reinterpret_cast<unsigned*>(a.bp)[r * kRecU32 + kMapOff + gg] = cur;
    }
}

// =========================================================================================== trace
// One warp per chain: e = state at the tile's last observation; the chain ends in state 0 (hmm.cpp:96).
// The lanes fetch 128 map words at a time (the next 128 are already in flight), the walk itself passes
// through them with one shuffle per tile, and every word is replaced by its e.
__global__ void __launch_bounds__(128)
viterbi_trace_kernel(ViterbiArgs a, int G)
{
    constexpr unsigned kFull = 0xffffffffu;
    const int wid = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wid >= a.n_samples * a.n_list) return;
    const int smp = wid % a.n_samples;                      // neighbouring warps: same chromosome, same length
    if (a.only_bad && a.seg_flags[1] == 0) return;
    const int chain = a.chain_list ? a.chain_list[wid / a.n_samples] : wid / a.n_samples;
    if (a.only_bad && a.seg_flags[seg_off_chain(chain)] == 0) return;
    const int grp = smp / G, gg = smp - grp * G;
    const ChainDesc cd = a.chains[chain];
    const int n_tiles = chain_tiles(cd);
    unsigned* __restrict__ tmap = reinterpret_cast<unsigned*>(a.bp) + record_base(a, chain, grp, n_tiles) * kRecU32 + kMapOff + gg;
    // 128 tiles per trip: lane l holds the maps of tiles top - 4l .. top - 4l - 3.  One trip costs a round trip to memory
    // (load the maps, store the states), and chromosome 1's 1,238 tiles took 39 of them with 32 tiles per trip: the latency
    // of the longest chain was the kernel's duration.
    constexpr int K = 4;
    unsigned e = 0;
    int top = n_tiles - 1;
    auto fetch = [&](int first, unsigned* w) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int t = first - (lane * K + k);
            w[k] = t >= 0 ? tmap[(int64_t)t * kRecU32] : 0u;
        }
    };
    unsigned nxt[K];
    fetch(top, nxt);
    while (top >= 0) {
        unsigned w[K], mine[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            w[k] = nxt[k];
            mine[k] = 0u;
        }
        if (top - 32 * K >= 0) fetch(top - 32 * K, nxt);
        const int n = top + 1 < 32 * K ? top + 1 : 32 * K;
        if (__all_sync(kFull, (w[0] | w[1] | w[2] | w[3]) == 0u)) {        // the usual case: every end state of these tiles came from state 0
            if (lane == 0) mine[0] = e;
            e = 0u;
        } else
            for (int l = 0; l * K < n; l++) {               // warp-uniform
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const unsigned wl = __shfl_sync(kFull, w[k], l);
                    if (l * K + k < n) {
                        if (lane == l) mine[k] = e;
                        e = (wl >> (4 * e)) & 0xFu;
                    }
                }
            }
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane * K + k < n) tmap[(int64_t)(top - (lane * K + k)) * kRecU32] = mine[k];
        top -= 32 * K;
    }
}

// =========================================================================================== expand
// One warp per (sample, chromosome), one lane per tile, 32 tiles at a time; state changes are rare and are
// replayed in ascending order through the reference's own scan (hmm.cpp:104-126, including its stale `start`).
template <int S>
__global__ void __launch_bounds__(128)
viterbi_expand_kernel(ViterbiArgs a)
{
    constexpr int G = 32 / S;
    constexpr unsigned kFull = 0xffffffffu;
    const int wid = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wid >= a.n_samples * a.n_list) return;
    const int smp = wid % a.n_samples;
    if (a.only_bad && a.seg_flags[1] == 0) return;
    const int chain = a.chain_list ? a.chain_list[wid / a.n_samples] : wid / a.n_samples;
    if (a.only_bad && a.seg_flags[seg_off_chain(chain)] == 0) return;
    const int grp = smp / G, gg = smp - grp * G;
    const ChainDesc cd = a.chains[chain];
    const int nobs = cd.nobs;
    const int64_t t_first = (cd.em_off + 1) >> 4;
    const int n_tiles = chain_tiles(cd);
    auto tile_i0 = [&](int t) -> int { return (int)(((t_first + t) << 4) - cd.em_off); };
    const uint2* __restrict__ bp = reinterpret_cast<const uint2*>(a.bp) + record_base(a, chain, grp, n_tiles) * kRecU2;
    const unsigned* __restrict__ tmap = reinterpret_cast<const unsigned*>(bp) + kMapOff;

    const int64_t tcell = (int64_t)smp * a.n_chains + chain;
    int32_t* __restrict__ slots = a.chain_calls + tcell * a.chain_call_cap * 4;
    int8_t* __restrict__ path = a.path + smp * a.path_stride + cd.out_off;
    const bool vec_path = (((reinterpret_cast<uintptr_t>(a.path) | (uintptr_t)a.path_stride) & 15) == 0) && cd.out_off == cd.em_off;
    int n_calls = 0, start = -1, nex = 0, run_start = 1;          // hmm.cpp:106, :108
    for (int base = 0; base < n_tiles; base += 32) {
        const int t = base + lane;
        const bool have = t < n_tiles;
        unsigned pw[4] = {0, 0, 0, 0};                      // states at the tile's 16 observations, one byte each
        unsigned bmask = 0, st = 0;
        if (have) {
            st = tmap[(int64_t)t * kRecU32 + gg] & 0xFu;
            uint2 w[S];
            w[0] = bp[(int64_t)t * kRecU2 + gg * S];
            // the usual case: the tile ends in state 0 and state 0's back-pointers are 0 throughout — sixteen zeros, no change
            // (and the other states' back-pointers are not even read)
            const bool quiet = st == 0u && (w[0].x | w[0].y) == 0u;
            if (!quiet) {
#pragma unroll
                for (int s = 1; s < S; s++) w[s] = bp[(int64_t)t * kRecU2 + gg * S + s];
            }
#pragma unroll
            for (int q = kTile - 1; q >= 0 && !quiet; q--) {
                pw[q >> 2] |= st << (8 * (q & 3));
                unsigned word = q < 8 ? w[0].x : w[0].y;
#pragma unroll
                for (int s = 1; s < S; s++) word = st == (unsigned)s ? (q < 8 ? w[s].x : w[s].y) : word;
                word = st == 7u ? 0u : word;                // reference reads out of bounds here; pinned to 0 like oracle.c
                const unsigned prev = (word >> (4 * (q & 7))) & 7u;
                bmask |= (prev != st ? 1u : 0u) << q;
                st = prev;
            }
            // st is now the state just before the tile
            const int i0 = tile_i0(t);
            unsigned ob[4];
#pragma unroll
            for (int r = 0; r < 4; r++) ob[r] = pw[r] | (((pw[r] + 0x01010101u) & 0x08080808u) * 31u);   // byte 7 -> 0xFF (-1)
            if (vec_path && i0 >= cd.out_first && i0 + kTile - 1 <= cd.out_last)
                __stcs(reinterpret_cast<uint4*>(path + i0), make_uint4(ob[0], ob[1], ob[2], ob[3]));
            else {
#pragma unroll
                for (int q = 0; q < kTile; q++)
                    if (i0 + q >= cd.out_first && i0 + q <= cd.out_last) path[i0 + q] = (int8_t)(ob[q >> 2] >> (8 * (q & 3)));
            }
            if (t == 0 && i0 == 1 && cd.out_first == 0) path[0] = (int8_t)(st == 7u ? -1 : (int)st);
        }
        unsigned ev = __ballot_sync(kFull, have && bmask != 0);
        while (ev) {                                        // warp-uniform replay, ascending tiles
            const int L = __ffs(ev) - 1;
            ev &= ev - 1;
            unsigned bm = __shfl_sync(kFull, bmask, L);
            const unsigned p0 = __shfl_sync(kFull, pw[0], L), p1 = __shfl_sync(kFull, pw[1], L);
            const unsigned p2 = __shfl_sync(kFull, pw[2], L), p3 = __shfl_sync(kFull, pw[3], L);
            const unsigned before = __shfl_sync(kFull, st, L);
            const int i0 = tile_i0(base + L);
            const unsigned long long plo = ((unsigned long long)p1 << 32) | p0, phi = ((unsigned long long)p3 << 32) | p2;
            while (bm) {
                const int q = __ffs(bm) - 1;
                bm &= bm - 1;
                const int i = i0 + q;                       // path[i-1] != path[i]   (hmm.cpp:110)
                const unsigned below = q == 0 ? before : (unsigned)(((q - 1) < 8 ? plo >> (8 * (q - 1)) : phi >> (8 * (q - 9))) & 0xFFu);
                const int below_state = below == 7u ? -1 : (int)below;
                if (below_state != 0) nex += i - run_start;            // hmm.cpp:124 over the run that ends at i-1
                run_start = i;
                const int current = i == 1 ? 0 : below_state;          // hmm.cpp:108, :125
                if (current == 0) start = i;                           // hmm.cpp:111
                else {                                                 // hmm.cpp:112-120
                    if (lane == 0 && n_calls < a.chain_call_cap) {
                        slots[4 * n_calls + 0] = start + 1 + cd.call_shift;
                        slots[4 * n_calls + 1] = i + cd.call_shift;
                        slots[4 * n_calls + 2] = current;
                        slots[4 * n_calls + 3] = nex;
                    }
                    n_calls++;
                    nex = 0;
                }
            }
        }
    }
    if (n_tiles == 0 && nobs >= 1 && cd.out_first == 0 && lane == 0) path[0] = 0;          // hmm.cpp:96
    if (lane == 0) a.chain_ncalls[tcell] = n_calls;
}

// =========================================================================================== compact
// One warp per sample: concatenate the per-chain call lists in chromosome order.
__global__ void viterbi_compact_kernel(ViterbiArgs a)
{
    const int sample = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (sample >= a.n_samples) return;
    int n = 0;
    for (int c = 0; c < a.n_chains; c++) {
        const int64_t t = (int64_t)sample * a.n_chains + c;
        const int m = a.chain_ncalls[t];
        const int have = m < a.chain_call_cap ? m : a.chain_call_cap;      // a chain that overflowed its scratch lost its LAST calls
        const int room = a.call_cap - n;
        const int copy = have < room ? have : (room > 0 ? room : 0);
        const int32_t* __restrict__ src = a.chain_calls + t * a.chain_call_cap * 4;
        int32_t* __restrict__ dst = a.calls + ((int64_t)sample * a.call_cap + n) * 4;
        for (int q = lane; q < copy * 4; q += 32) dst[q] = src[q];
        n += m;                 // the count stays honest; > call_cap signals truncation to the host
    }
    if (lane == 0) a.ncalls[sample] = n;
}

size_t viterbi_smem_bytes(int S, int W)
{
    return (size_t)W * (ring_stages(S, W) * (stage_bytes(S) + 16) + 2 * (32 / S + 1) * lt_jstride(S) * 8);
}
// 4 or 8 sweep warps per CTA: whichever finishes the batch sooner under the measured per-step costs (see kTile comment).
// Work items are indivisible, so the finishing time is the busiest warp's load under the real placement
// (viterbi_schedule), not the average load: 640 equal items on 592 warps take two rounds, on 1184 warps one.
int viterbi_pick_warps(const int32_t* chain_nobs, int n_chains, int groups, int n_sms)
{
    if (n_chains < 1 || groups < 1 || n_sms < 1) return 4;
    double t[2];
    for (int v = 0; v < 2; v++) {
        const int W = v ? 8 : 4;
        std::vector<int32_t> begin, items;
        viterbi_schedule(chain_nobs, n_chains, groups, n_sms, W, begin, items);
        int64_t busiest = 0;
        for (int s = 0; s + 1 < (int)begin.size(); s++) {
            int64_t load = 0;
            for (int q = begin[s]; q < begin[s + 1]; q++) load += chain_nobs[items[2 * q]];
            if (load > busiest) busiest = load;
        }
        t[v] = (v ? 185.0 : 140.0) * (double)busiest;
    }
    return t[0] <= t[1] ? 4 : 8;
}
int viterbi_lt_pitch(int S) { return lt_pitch(S); }
int viterbi_tile() { return kTile; }
size_t viterbi_record_bytes() { return (size_t)kRecU2 * 8; }

template <int S, int W>
static void launch_sweep(const ViterbiArgs& a, cudaStream_t st)
{
    const size_t smem = viterbi_smem_bytes(S, W);
    static PerDevice configured;
    if (configured.raise(smem)) cudaFuncSetAttribute(viterbi_sweep_kernel<S, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    viterbi_sweep_kernel<S, W><<<a.n_slots / W, (W + 1) * 32, smem, st>>>(a, *reinterpret_cast<const CUtensorMap*>(a.ll_map));
}

static const char* const kPassNames[4] = {"viterbi_sweep", "viterbi_tilemap", "viterbi_trace", "viterbi_expand"};
static const char* const kRepairNames[4] = {"repair_sweep", "repair_tilemap", "repair_trace", "repair_expand"};
template <int S>
static void launch_all(const ViterbiArgs& a, cudaStream_t st)
{
    constexpr int G = 32 / S;
    // (the repair pass of a segmented sweep is timed under names of its own: normally four launches that find nothing to do)
    const char* const* nm = a.only_bad ? kRepairNames : kPassNames;
    prof_mark(nm[0], st);
    if (a.tpc) launch_viterbi_tpc_sweep(a, st);
    else switch (a.warps_per_cta) {
        case 1: launch_sweep<S, 1>(a, st); break;          // 1, 2: experiments (EDB200_CRIT_WARPS), see DESIGN.md "what comes next"
        case 2: launch_sweep<S, 2>(a, st); break;
        case 4: launch_sweep<S, 4>(a, st); break;
        default: launch_sweep<S, 8>(a, st); break;
    }
    prof_mark(nm[1], st);
    if (a.flat_records > 0 && !a.only_bad)
        viterbi_tilemap_kernel<S><<<(unsigned)((a.flat_records * G + 255) / 256), 256, 0, st>>>(a);
    else {
        ViterbiArgs b = a;
        b.flat_records = 0;
        const int64_t map_threads = (int64_t)a.max_list_tiles * a.groups * G;
        const unsigned bx = (unsigned)((map_threads + 255) / 256);
        if (map_threads > 0) viterbi_tilemap_kernel<S><<<dim3(a.only_bad ? std::min(bx, 16u) : bx, (unsigned)a.n_list), 256, 0, st>>>(b);
    }
    const int chains = a.n_samples * a.n_list;
    prof_mark(nm[2], st);
    viterbi_trace_kernel<<<(chains + 3) / 4, 128, 0, st>>>(a, G);
    prof_mark(nm[3], st);
    viterbi_expand_kernel<S><<<(chains + 3) / 4, 128, 0, st>>>(a);
}

int launch_viterbi(const ViterbiArgs& a, cudaStream_t st)
{
    if (a.n_list == 0 || a.n_samples == 0) return 0;
    switch (a.n_states) {
        case 2: launch_all<2>(a, st); break;
        case 3: launch_all<3>(a, st); break;
        case 4: launch_all<4>(a, st); break;
        case 5: launch_all<5>(a, st); break;
        case 6: launch_all<6>(a, st); break;
        case 7: launch_all<7>(a, st); break;
        default: return 0;
    }
    prof_mark(nullptr, st);
    return 4;
}

int launch_viterbi_compact(const ViterbiArgs& a, cudaStream_t st)
{
    if (a.n_chains == 0 || a.n_samples == 0) return 0;
    prof_mark("viterbi_compact", st);
    viterbi_compact_kernel<<<(a.n_samples + 3) / 4, 128, 0, st>>>(a);
    prof_mark(nullptr, st);
    return 1;
}

}  // namespace edb
