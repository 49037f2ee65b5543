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
// a chain exchange V[i-1][k] with warp shuffles.  Back-pointers leave the warp as three ballots per
// step (one per bit), i.e. 4 bytes per chain·observation.
#include "kernels.cuh"

namespace edb {

template <int S>
__global__ void __launch_bounds__(128)
viterbi_forward_kernel(ViterbiArgs a, int groups_per_chain)
{
    constexpr int G = 32 / S;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int chain = warp_global / groups_per_chain;
    const int grp = warp_global - chain * groups_per_chain;
    if (chain >= a.n_chains) return;

    int g = lane / S;
    const int j = lane - g * S;
    const bool lane_ok = g < G;
    if (!lane_ok) g = G - 1;
    int sample = grp * G + g;
    const bool valid = lane_ok && sample < a.n_samples;
    if (sample >= a.n_samples) sample = a.n_samples - 1;

    const ChainDesc cd = a.chains[chain];
    const double* __restrict__ em_row = a.ll + sample * a.ll_sample_stride + a.perm[j] * a.ll_state_stride + cd.em_off;
    const double* __restrict__ lt_row = a.lt + cd.lt_row0 * (S * S) + j * S;
    uint32_t* __restrict__ bp = a.bp + sample * a.bp_stride + cd.lt_row0;
    const double ninf = -HUGE_VAL;
    const double tail = j == 0 ? 0.0 : a.tail_other;
    const unsigned fieldmask = (1u << S) - 1u;
    const int src0 = g * S;

    double V = j == 0 ? 0.0 : ninf;                         // hmm.cpp:46-52
    for (int i = 1; i < cd.nobs; i++) {
        const double em = i <= cd.n_em ? em_row[i] : tail;
        const double* __restrict__ lti = lt_row + (int64_t)i * (S * S);
        double best = ninf;
        int arg = 7;                                        // 7 encodes "from = -1" (hmm.cpp:60)
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double vk = __shfl_sync(0xffffffffu, V, src0 + k);
            const double cand = __dadd_rn(__dadd_rn(em, vk), lti[k]);   // hmm.cpp:79
            if (cand > best) { best = cand; arg = k; }                  // hmm.cpp:81
        }
        if (em == ninf) arg = 0;                            // hmm.cpp:87
        V = best;
        const unsigned b0 = __ballot_sync(0xffffffffu, arg & 1);
        const unsigned b1 = __ballot_sync(0xffffffffu, arg & 2);
        const unsigned b2 = __ballot_sync(0xffffffffu, arg & 4);
        if (j == 0 && valid)
            bp[i] = ((b0 >> src0) & fieldmask) | (((b1 >> src0) & fieldmask) << 8) | (((b2 >> src0) & fieldmask) << 16);
    }
}

// One thread per (sample, chain): traceback (hmm.cpp:95-100), then the reference's forward segment
// scan (hmm.cpp:104-126) over the states it just wrote.
__global__ void __launch_bounds__(128)
viterbi_traceback_kernel(ViterbiArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_samples * a.n_chains) return;
    const int chain = t % a.n_chains, sample = t / a.n_chains;
    const ChainDesc cd = a.chains[chain];
    uint32_t* __restrict__ bp = a.bp + sample * a.bp_stride + cd.lt_row0;
    int8_t* __restrict__ path = a.path + sample * a.path_stride + cd.out_off;

    int st = 0;                                             // hmm.cpp:96
    for (int i = cd.nobs - 1; i >= 1; i--) {
        const uint32_t w = bp[i];
        bp[i] = (uint32_t)st;                               // slot reused for the decoded state
        int prev;
        if (st < 0) prev = 0;                               // reference reads out of bounds here; pinned to 0 like oracle.c
        else {
            prev = ((w >> st) & 1u) | (((w >> (8 + st)) & 1u) << 1) | (((w >> (16 + st)) & 1u) << 2);
            if (prev == 7) prev = -1;
        }
        st = prev;
    }
    if (cd.nobs > 0) bp[0] = (uint32_t)st;

    int32_t* __restrict__ calls = a.chain_calls + ((int64_t)t * a.chain_call_cap) * 4;
    int n = 0, current = 0, start = -1, nex = 0;
    int prev_state = cd.nobs > 0 ? (int)bp[0] : 0;
    if (cd.out_first == 0 && cd.nobs > 0) path[0] = (int8_t)prev_state;
    for (int i = 1; i < cd.nobs; i++) {
        const int cur = (int)bp[i];
        if (prev_state != cur) {
            if (current == 0) start = i;
            else {
                if (n < a.chain_call_cap) {
                    calls[4 * n + 0] = start + 1 + cd.call_shift;
                    calls[4 * n + 1] = i + cd.call_shift;
                    calls[4 * n + 2] = current;
                    calls[4 * n + 3] = nex;
                }
                n++;
                nex = 0;
            }
        }
        if (cur != 0) nex++;
        current = cur;
        prev_state = cur;
        if (i >= cd.out_first && i <= cd.out_last) path[i] = (int8_t)cur;
    }
    a.chain_ncalls[t] = n;
}

// One thread per sample: concatenate the per-chain call lists in chromosome order.
__global__ void viterbi_compact_kernel(ViterbiArgs a)
{
    const int sample = blockIdx.x * blockDim.x + threadIdx.x;
    if (sample >= a.n_samples) return;
    int n = 0;
    for (int c = 0; c < a.n_chains; c++) {
        const int64_t t = (int64_t)sample * a.n_chains + c;
        const int m = a.chain_ncalls[t];
        const int have = m < a.chain_call_cap ? m : a.chain_call_cap;
        const int32_t* src = a.chain_calls + t * a.chain_call_cap * 4;
        for (int q = 0; q < have; q++, n++)
            if (n < a.call_cap)
                for (int f = 0; f < 4; f++) a.calls[((int64_t)sample * a.call_cap + n) * 4 + f] = src[4 * q + f];
        n += m - have;
    }
    a.ncalls[sample] = n;      // > call_cap signals truncation to the host
}

void launch_viterbi(const ViterbiArgs& a, cudaStream_t st)
{
    if (a.n_chains == 0 || a.n_samples == 0) return;
    const int S = a.n_states;
    const int G = 32 / S;
    const int groups = (a.n_samples + G - 1) / G;
    const int warps = groups * a.n_chains;
    const int blocks = (warps + 3) / 4;
    switch (S) {
        case 2: viterbi_forward_kernel<2><<<blocks, 128, 0, st>>>(a, groups); break;
        case 3: viterbi_forward_kernel<3><<<blocks, 128, 0, st>>>(a, groups); break;
        case 4: viterbi_forward_kernel<4><<<blocks, 128, 0, st>>>(a, groups); break;
        case 5: viterbi_forward_kernel<5><<<blocks, 128, 0, st>>>(a, groups); break;
        case 6: viterbi_forward_kernel<6><<<blocks, 128, 0, st>>>(a, groups); break;
        case 7: viterbi_forward_kernel<7><<<blocks, 128, 0, st>>>(a, groups); break;
        default: return;
    }
    const int nt = a.n_samples * a.n_chains;
    viterbi_traceback_kernel<<<(nt + 127) / 128, 128, 0, st>>>(a);
    viterbi_compact_kernel<<<(a.n_samples + 127) / 128, 128, 0, st>>>(a);
}

}  // namespace edb
