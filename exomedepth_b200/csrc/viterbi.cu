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
// Mapping: one lane per (chain, destination state).  A warp carries G = 32/S chains of the SAME
// chromosome (so every lane runs the same number of steps) from G consecutive samples; the S lanes of
// a chain exchange V[i-1][k] with warp shuffles.  Each warp is its own pipeline:
//   * the shared log-transition rows stream through a per-warp shared-memory ring with TMA
//     (cp.async.bulk + mbarrier complete_tx), 16 observations per tile;
//   * every lane prefetches its own emission row one tile ahead with 128-bit loads (one full 128-byte
//     line per lane and tile) into registers, so the sequential recurrence never waits on memory;
//   * every lane packs its 16 back-pointers of a tile into 64 bits (4 bits each) and stores them once
//     per tile: 256 bytes per warp·tile, coalesced;
//   * the same warp then walks them backwards (traceback, hmm.cpp:95-100) and produces the reference's
//     call table during that walk.
#include "kernels.cuh"

namespace edb {

constexpr int kTile = 16;          // observations per tile (one 128-byte line of an emission row)
constexpr int kStages = 4;         // TMA ring depth for the transition rows
constexpr int kWarpsPerCta = 8;

__host__ __device__ constexpr int lt_pitch(int S) { return S * S + ((S * S) & 1); }   // doubles per row, 16-byte multiple

// ---- PTX helpers (mbarrier + 1-D bulk TMA) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- call-table bookkeeping during the backward walk ------------------------------------------------
// The reference scans forward (hmm.cpp:104-126): at every state change i it either records start = i (previous
// state normal) or emits (start+1, i, previous state, run length).  Walking backwards the same calls appear
// last-first; `start` of a call is the first observation of the enclosing block of non-normal states, known
// only when the walk reaches it, so the calls of the open block are patched then.
struct CallWriter {
    int32_t* slots;     // per-chain scratch, filled from the end
    int cap, n;         // n = calls written so far
    int open_from;      // first (lowest slot index) call of the block still waiting for its start
    int pending_end;    // end observation of the newest call, waiting for its run length
    int shift;
    __device__ void emit(int end_i, int type)
    {
        n++;
        const int s = cap - n;
        if (s >= 0) {
            slots[4 * s + 1] = end_i + shift;
            slots[4 * s + 2] = type;
        }
        pending_end = end_i;
    }
    __device__ void run_starts(int i2)      // the newest call's run is [i2 .. end-1]
    {
        const int s = cap - n;
        if (n > 0 && s >= 0 && pending_end >= 0) slots[4 * s + 3] = pending_end - i2;
        pending_end = -1;
    }
    __device__ void block_starts(int i3)    // start = i3 for every call of the open block
    {
        for (int q = open_from; q < n; q++) {
            const int s = cap - 1 - q;
            if (s >= 0) slots[4 * s + 0] = i3 + 1 + shift;
        }
        open_from = n;
    }
};

// ---- one step of the recurrence for this lane's destination state ---------------------------------
// cand_k = (em + V[k]) + lt[k]  (hmm.cpp:79), winner = FIRST maximum (strict '>' of hmm.cpp:81).
// The maximum is taken with an order-preserving tournament: a node keeps its left entry unless the right
// one is strictly greater, which selects exactly the entry the reference's sequential scan selects.
// NaN never reaches the tournament: NaN transition terms are stored as -Inf in the device copy of the
// table (a NaN candidate and a -Inf candidate are both "never selected", hmm.cpp:81) and a NaN emission
// is replaced by -Inf (all candidates lose, from stays -1, V stays -Inf, as in the reference).
// SPECIAL = false is the hot variant for tiles whose emissions are all finite.
template <int S, bool SPECIAL>
__device__ __forceinline__ unsigned viterbi_step(double em, const double* __restrict__ lt_qj, int src0, double& V)
{
    const double ninf = -HUGE_VAL;
    const double em_s = (SPECIAL && em != em) ? ninf : em;
    double c[S];
    int id[S];
#pragma unroll
    for (int k = 0; k < S; k++) {
        const double vk = __shfl_sync(0xffffffffu, V, src0 + k);
        c[k] = __dadd_rn(__dadd_rn(em_s, vk), lt_qj[k]);
        id[k] = k;
    }
#pragma unroll
    for (int n = S; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int p = 0; p + 1 < n; p += 2) {
            const bool right = c[p + 1] > c[p];
            c[p / 2] = right ? c[p + 1] : c[p];
            id[p / 2] = right ? id[p + 1] : id[p];
        }
        if (n & 1) { c[n / 2] = c[n - 1]; id[n / 2] = id[n - 1]; }
    }
    V = c[0];
    unsigned arg = c[0] > ninf ? (unsigned)id[0] : 7u;      // 7 encodes "from = -1" (hmm.cpp:60)
    if (SPECIAL && em == ninf) arg = 0u;                    // hmm.cpp:87
    return arg;
}

// which sorted warp slot does warp `w` of CTA `c` run?  Slots are ordered longest chain first; lanes 0-3 of a
// CTA take the front of the list and lanes 4-7 the back, so each SM sub-partition (warp id mod 4) hosts one
// long and one short chain.
__device__ __forceinline__ int warp_slot(int c, int w, int n_warps)
{
    const int half = (n_warps + 1) / 2;
    if (w < 4) {
        const int f = 4 * c + w;
        return f < half ? f : -1;
    }
    const int b = n_warps - 1 - (4 * c + (w - 4));
    return b >= half ? b : -1;
}

template <int S>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
viterbi_chain_kernel(ViterbiArgs a, int groups_per_chain)
{
    constexpr int G = 32 / S;
    constexpr int LTP = lt_pitch(S);
    constexpr unsigned kTileBytes = kTile * LTP * 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ring = reinterpret_cast<double*>(smem) + (size_t)warp * kStages * kTile * LTP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarpsPerCta * kStages * kTileBytes) + warp * kStages;

    const int warp_global = warp_slot(blockIdx.x, warp, a.n_chains * groups_per_chain);
    if (warp_global < 0) return;                            // each warp is an independent pipeline: no CTA barrier below
    const int chain = a.order ? a.order[warp_global / groups_per_chain] : warp_global / groups_per_chain;
    const int grp = warp_global % groups_per_chain;

    int g = lane / S;
    const int j = lane - g * S;
    const bool lane_ok = g < G;
    if (!lane_ok) g = G - 1;
    int sample = grp * G + g;
    const bool valid = lane_ok && sample < a.n_samples;
    if (sample >= a.n_samples) sample = a.n_samples - 1;
    const int src0 = g * S;

    const ChainDesc cd = a.chains[chain];
    const int nobs = cd.nobs;
    // tiles follow the 128-byte lines of the emission rows: tile t covers observations i with
    // (em_off + i) / 16 == t_first + t
    const int64_t t_first = (cd.em_off + 1) >> 4;
    const int n_tiles = nobs > 1 ? (int)(((cd.em_off + nobs - 1) >> 4) - t_first + 1) : 0;
    const double* __restrict__ em_row = a.ll + sample * a.ll_sample_stride + a.perm[j] * a.ll_state_stride;
    const double* __restrict__ lt_base = a.lt + cd.lt_row0 * LTP;
    uint2* __restrict__ bp = reinterpret_cast<uint2*>(a.bp) +
                             ((int64_t)a.bp_tile_base[chain] * groups_per_chain + (int64_t)grp * n_tiles) * 32;

    // ---------------------------------------------------------------- TMA ring for the transition rows
    auto tile_i0 = [&](int t) -> int { return (int)(((t_first + t) << 4) - cd.em_off); };   // first obs of tile (may be < 1)
    auto issue_lt = [&](int t) {
        const int st = t % kStages;
        const int i0 = tile_i0(t);
        const int r0 = i0 < 0 ? 0 : i0;                 // rows before the chain's first row are never used
        const unsigned bytes = (unsigned)(i0 + kTile - r0) * LTP * 8;
        mbar_expect_tx(&bars[st], bytes);
        tma_load_1d(ring + ((size_t)st * kTile + (r0 - i0)) * LTP, lt_base + (int64_t)r0 * LTP, bytes, &bars[st]);
    };
    if (lane == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
        for (int t = 0; t < kStages && t < n_tiles; t++) issue_lt(t);

    // ---------------------------------------------------------------- emission prefetch (registers)
    double em_nxt[kTile];
#pragma unroll
    for (int q = 0; q < kTile; q++) em_nxt[q] = 0.0;
    auto load_em = [&](int t) {
        // tiles holding only the dummy last observation have no emission row behind them
        if (tile_i0(t) <= cd.n_em) {
            const double2* p = reinterpret_cast<const double2*>(em_row + ((t_first + t) << 4));
#pragma unroll
            for (int q = 0; q < kTile / 2; q++) {
                const double2 v = __ldcs(p + q);          // streamed once: evict-first
                em_nxt[2 * q] = v.x;
                em_nxt[2 * q + 1] = v.y;
            }
        }
    };
    if (n_tiles > 0) load_em(0);

    const double tail = j == 0 ? 0.0 : a.tail_other;
    double V = j == 0 ? 0.0 : -HUGE_VAL;                    // hmm.cpp:46-52

    for (int t = 0; t < n_tiles; t++) {
        double em_cur[kTile];
        bool special = false;
#pragma unroll
        for (int q = 0; q < kTile; q++) {
            em_cur[q] = em_nxt[q];
            special |= !(em_cur[q] > -HUGE_VAL);           // NaN or -Inf somewhere in the tile
        }
        if (t + 1 < n_tiles) load_em(t + 1);
        const int st = t % kStages;
        mbar_wait(&bars[st], (t / kStages) & 1);
        const double* __restrict__ ltt = ring + (size_t)st * kTile * LTP + j * S;
        const int i0 = tile_i0(t);
        unsigned lo = 0, hi = 0;                            // 16 back-pointers of this lane, 4 bits each
        const bool whole = i0 >= 1 && i0 + kTile - 1 <= cd.n_em;   // tile entirely inside the real observations
        if (whole && !__any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int q = 0; q < kTile; q++) {
                const unsigned arg = viterbi_step<S, false>(em_cur[q], ltt + q * LTP, src0, V);
                if (q < 8) lo |= arg << (4 * q);
                else hi |= arg << (4 * (q - 8));
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < kTile; q++) {
                const int i = i0 + q;
                if (i >= 1 && i < nobs) {                   // warp-uniform
                    double em = tail;
#pragma unroll
                    for (int r = 0; r < kTile; r++) em = (r == q && i <= cd.n_em) ? em_cur[r] : em;
                    const unsigned arg = viterbi_step<S, true>(em, ltt + q * LTP, src0, V);
                    if (q < 8) lo |= arg << (4 * q);
                    else hi |= arg << (4 * (q - 8));
                }
            }
        }
        bp[(int64_t)t * 32 + lane] = make_uint2(lo, hi);
        __syncwarp();
        if (lane == 0 && t + kStages < n_tiles) issue_lt(t + kStages);
    }
    __syncwarp();
    __threadfence_block();

    // ---------------------------------------------------------------- traceback + call table (leader lanes)
    if (j != 0 || !valid) return;
    const int64_t tcell = (int64_t)sample * a.n_chains + chain;
    CallWriter cw;
    cw.slots = a.chain_calls + tcell * a.chain_call_cap * 4;
    cw.cap = a.chain_call_cap;
    cw.n = 0;
    cw.open_from = 0;
    cw.pending_end = -1;
    cw.shift = cd.call_shift;
    // path index of observation i is out_off + i, and out_off == em_off in both framings, so a tile's 16
    // observations are one aligned 16-byte group of the path row
    int8_t* __restrict__ path = a.path + sample * a.path_stride + cd.out_off;
    const bool vec_path = (((reinterpret_cast<uintptr_t>(a.path) | (uintptr_t)a.path_stride) & 15) == 0) && cd.out_off == cd.em_off;

    // a state change between observation i-1 (state `below`) and i (hmm.cpp:110), met while walking down
    auto boundary = [&](int i, int below) {
        const int prev_state = below == 7 ? -1 : below;
        cw.run_starts(i);                                   // a run that was open above starts at i
        const int current = i == 1 ? 0 : prev_state;        // hmm.cpp:108 starts with current = 0
        if (current == 0) cw.block_starts(i);               // hmm.cpp:111
        else cw.emit(i, current);                           // hmm.cpp:112-120
    };

    unsigned st = 0;                                        // state at observation nobs-1 (hmm.cpp:96)
    uint2 wn[S];
    auto load_bp = [&](int t) {
#pragma unroll
        for (int s = 0; s < S; s++) wn[s] = bp[(int64_t)t * 32 + src0 + s];
    };
    if (n_tiles > 0) load_bp(n_tiles - 1);
    for (int t = n_tiles - 1; t >= 0; t--) {
        uint2 w[S];
#pragma unroll
        for (int s = 0; s < S; s++) w[s] = wn[s];
        if (t > 0) load_bp(t - 1);
        const int i0 = tile_i0(t);
        const bool whole = i0 >= 1 && i0 + kTile - 1 <= cd.out_last && i0 + kTile - 1 < nobs && i0 >= cd.out_first;
        if (whole) {
            unsigned pw[4] = {0, 0, 0, 0};                  // states at observations i0 .. i0+15, one byte each
            unsigned bmask = 0;
            const unsigned st_top = st;
#pragma unroll
            for (int q = kTile - 1; q >= 0; q--) {
                pw[q >> 2] |= st << (8 * (q & 3));
                unsigned word = q < 8 ? w[0].x : w[0].y;
#pragma unroll
                for (int s = 1; s < S; s++) word = st == (unsigned)s ? (q < 8 ? w[s].x : w[s].y) : word;
                word = st == 7u ? 0u : word;                // reference reads out of bounds here; pinned to 0 like oracle.c
                const unsigned prev = (word >> (4 * (q & 7))) & 7u;
                bmask |= (prev != st ? 1u : 0u) << q;
                st = prev;
            }
            if (bmask) {                                    // rare: replay the tile's state changes top-down
                unsigned above = st_top;
                (void)above;
                for (int q = kTile - 1; q >= 0; q--) {
                    if ((bmask >> q) & 1u) {
                        const unsigned below = q > 0 ? (pw[(q - 1) >> 2] >> (8 * ((q - 1) & 3))) & 0xffu : st;
                        boundary(i0 + q, (int)below);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {                   // byte 7 -> 0xFF (-1)
                const unsigned m = (pw[r] + 0x01010101u) & 0x08080808u;
                pw[r] |= m * 31u;
            }
            if (vec_path) *reinterpret_cast<uint4*>(path + i0) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
            else {
#pragma unroll
                for (int q = 0; q < kTile; q++) path[i0 + q] = (int8_t)(pw[q >> 2] >> (8 * (q & 3)));
            }
        } else {
#pragma unroll 1
            for (int q = kTile - 1; q >= 0; q--) {
                const int i = i0 + q;
                if (i >= 1 && i < nobs) {
                    if (i >= cd.out_first && i <= cd.out_last) path[i] = (int8_t)(st == 7u ? -1 : (int)st);
                    unsigned word = 0;
#pragma unroll
                    for (int s = 0; s < S; s++) word = st == (unsigned)s ? (q < 8 ? w[s].x : w[s].y) : word;
                    const unsigned prev = st == 7u ? 0u : (word >> (4 * (q & 7))) & 7u;
                    if (prev != st) boundary(i, (int)prev);
                    st = prev;
                }
            }
        }
    }
    // observation 0 reached
    if (nobs >= 1 && cd.out_first == 0) path[0] = (int8_t)(st == 7u ? -1 : (int)st);
    cw.run_starts(1);                                       // a run still open extends to the chain start (hmm.cpp:124 counts from obs 1)
    cw.block_starts(-1);                                    // start keeps its initial -1 (hmm.cpp:106)
    a.chain_ncalls[tcell] = cw.n;
}

// One thread per sample: concatenate the per-chain call lists (stored last-first at the end of each
// chain's scratch) in chromosome order.
__global__ void viterbi_compact_kernel(ViterbiArgs a)
{
    const int sample = blockIdx.x * blockDim.x + threadIdx.x;
    if (sample >= a.n_samples) return;
    int n = 0;
    for (int c = 0; c < a.n_chains; c++) {
        const int64_t t = (int64_t)sample * a.n_chains + c;
        const int m = a.chain_ncalls[t];
        const int have = m < a.chain_call_cap ? m : a.chain_call_cap;
        const int32_t* src = a.chain_calls + (t * a.chain_call_cap + (a.chain_call_cap - have)) * 4;
        // when the chain overflowed its scratch, the EARLIEST calls were dropped; keep the count honest
        n += m - have;
        for (int q = 0; q < have; q++, n++)
            if (n < a.call_cap)
                for (int f = 0; f < 4; f++) a.calls[((int64_t)sample * a.call_cap + n) * 4 + f] = src[4 * q + f];
    }
    a.ncalls[sample] = n;      // > call_cap signals truncation to the host
}

size_t viterbi_smem_bytes(int S) { return (size_t)kWarpsPerCta * kStages * (kTile * lt_pitch(S) * 8 + 8); }
int viterbi_lt_pitch(int S) { return lt_pitch(S); }
int viterbi_tile() { return kTile; }

template <int S>
static void launch_chain(const ViterbiArgs& a, int blocks, int groups, cudaStream_t st)
{
    const size_t smem = viterbi_smem_bytes(S);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(viterbi_chain_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    viterbi_chain_kernel<S><<<blocks, kWarpsPerCta * 32, smem, st>>>(a, groups);
}

void launch_viterbi(const ViterbiArgs& a, cudaStream_t st)
{
    if (a.n_chains == 0 || a.n_samples == 0) return;
    const int S = a.n_states;
    const int G = 32 / S;
    const int groups = (a.n_samples + G - 1) / G;
    const int warps = groups * a.n_chains;
    const int blocks = ((warps + 1) / 2 + 3) / 4;          // see warp_slot(): 4 front + 4 back slots per CTA
    switch (S) {
        case 2: launch_chain<2>(a, blocks, groups, st); break;
        case 3: launch_chain<3>(a, blocks, groups, st); break;
        case 4: launch_chain<4>(a, blocks, groups, st); break;
        case 5: launch_chain<5>(a, blocks, groups, st); break;
        case 6: launch_chain<6>(a, blocks, groups, st); break;
        case 7: launch_chain<7>(a, blocks, groups, st); break;
        default: return;
    }
    viterbi_compact_kernel<<<(a.n_samples + 127) / 128, 128, 0, st>>>(a);
}

}  // namespace edb
