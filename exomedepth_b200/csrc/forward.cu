// forward.cu — HMM forward pass and transition-probability grid (sm_100a, FP64).
//
// EXTENSION: the reference has no forward pass and no transition-probability estimate (its
// transition.probability is a user constant, R/class_definition.R:261).  The
// definition followed here is oracle/oracle.c:edo_forward_loglik — the same distance-dependent transition
// model as the Viterbi sweep (src/hmm.cpp:62-76), alpha_0 = (1, 0, ..), forced end in the normal state,
// transition terms that are zero, negative or NaN contribute nothing (the strict '>' of hmm.cpp:81 skips
// them in the Viterbi sweep) — evaluated for a grid of transition probabilities at once; the maximiser
// over the grid is the per-sample MLE.  Parity is UNPINNED (no reference code); tests compare with the
// oracle port to 1e-10 relative.
//
// One thread per (sample, chromosome, grid point), the S forward weights in registers, rescaled every
// observation (linear domain: 5 exp + 1 log per step instead of the 25 exp + 5 log of the log-space
// recurrence).  Lanes of a warp are consecutive grid points of the same chain, so the emission loads are
// warp-wide broadcasts.  The per-chromosome results are summed in chromosome order by a second kernel, so
// the totals do not depend on scheduling.
#include "kernels.cuh"

namespace edb {

template <int S>
__global__ void __launch_bounds__(128)
forward_chain_kernel(ForwardArgs a)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)a.n_chains * a.n_samples * a.n_grid;
    if (idx >= total) return;
    const int gi = (int)(idx % a.n_grid);
    const int64_t rest = idx / a.n_grid;
    const int smp = (int)(rest % a.n_samples), chain = (int)(rest / a.n_samples);
    const ChainDesc cd = a.chains[chain];

    // T[k + S*j] = P(k -> j) for this grid point
    double T[S * S];
#pragma unroll
    for (int q = 0; q < S * S; q++) T[q] = a.T_grid[(int64_t)gi * S * S + q];

    const double* __restrict__ em_base = a.ll + smp * a.ll_sample_stride + cd.em_off;
    const double* __restrict__ decay = a.decay + cd.lt_row0;
    double al[S];
    al[0] = 1.0;
#pragma unroll
    for (int j = 1; j < S; j++) al[j] = 0.0;
    double logl = 0.0;
    bool dead = false;

    for (int i = 1; i < cd.nobs; i++) {
        double em[S];
        if (i <= cd.n_em) {
#pragma unroll
            for (int j = 0; j < S; j++) em[j] = em_base[a.perm[j] * a.ll_state_stride + i];
        } else {
            em[0] = 0.0;
#pragma unroll
            for (int j = 1; j < S; j++) em[j] = a.tail_other;
        }
        double m = -HUGE_VAL;
#pragma unroll
        for (int j = 0; j < S; j++) {
            if (!(em[j] == em[j])) em[j] = -HUGE_VAL;       // a NaN emission cannot be reached
            m = fmax(m, em[j]);
        }
        if (m == -HUGE_VAL) { dead = true; break; }
        const double d = decay[i], omd = 1.0 - d;
        double nw[S], s = 0.0;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const double t0 = T[j * S];
            double acc = al[0] * (t0 > 0.0 ? t0 : 0.0);
#pragma unroll
            for (int k = 1; k < S; k++) {
                const double t = d * T[j * S + k] + omd * t0;             // hmm.cpp:75-76
                acc = fma(al[k], t > 0.0 ? t : 0.0, acc);
            }
            nw[j] = acc * exp(em[j] - m);
            s += nw[j];
        }
        if (!(s > 0.0)) { dead = true; break; }
        const double r = 1.0 / s;
#pragma unroll
        for (int j = 0; j < S; j++) al[j] = nw[j] * r;
        logl += m + log(s);
    }
    const double out = dead ? -HUGE_VAL : logl + log(al[0]);             // forced end in the normal state (hmm.cpp:96)
    a.chain_loglik[((int64_t)smp * a.n_chains + chain) * a.n_grid + gi] = out;
}

// loglik[sample][grid] = sum over chromosomes, in chromosome order; best[sample] = first maximiser
__global__ void forward_reduce_kernel(ForwardArgs a)
{
    const int smp = blockIdx.x * blockDim.x + threadIdx.x;
    if (smp >= a.n_samples) return;
    int best = 0;
    double best_v = -HUGE_VAL;
    for (int gi = 0; gi < a.n_grid; gi++) {
        double s = 0.0;
        for (int c = 0; c < a.n_chains; c++) s += a.chain_loglik[((int64_t)smp * a.n_chains + c) * a.n_grid + gi];
        a.loglik[(int64_t)smp * a.n_grid + gi] = s;
        if (s > best_v) { best_v = s; best = gi; }
    }
    if (a.best) a.best[smp] = best;
}

int launch_forward(const ForwardArgs& a, cudaStream_t st)
{
    const int64_t total = (int64_t)a.n_chains * a.n_samples * a.n_grid;
    if (total == 0) return 0;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    prof_mark("forward_chain", st);
    switch (a.n_states) {
        case 2: forward_chain_kernel<2><<<blocks, 128, 0, st>>>(a); break;
        case 3: forward_chain_kernel<3><<<blocks, 128, 0, st>>>(a); break;
        case 4: forward_chain_kernel<4><<<blocks, 128, 0, st>>>(a); break;
        case 5: forward_chain_kernel<5><<<blocks, 128, 0, st>>>(a); break;
        case 6: forward_chain_kernel<6><<<blocks, 128, 0, st>>>(a); break;
        case 7: forward_chain_kernel<7><<<blocks, 128, 0, st>>>(a); break;
        default: return 0;
    }
    prof_mark("forward_reduce", st);
    forward_reduce_kernel<<<(a.n_samples + 127) / 128, 128, 0, st>>>(a);
    prof_mark(nullptr, st);
    return 2;
}

}  // namespace edb
