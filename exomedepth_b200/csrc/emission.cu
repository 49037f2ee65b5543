// emission.cu — beta-binomial emission log-likelihood kernels (sm_100a, FP64).
//
// Replaces get_loglike_matrix / myprob of the reference (src/CNV_estimate.cpp:44-50, 52-85) and the
// vendored lnbeta chain under it (src/beta.c, src/VP_gamma.c, src/VP_log.c).
//
//   emission_bins_kernel    reference-API shape: one sample, per-bin phi / expected vectors.
//   emission_direct_kernel  cohort shape: per-sample scalar phi / expected, constants hoisted per
//                           (sample, state), lgamma differences in registers (FP64-pipe bound).
//   emission_table_kernel   cohort shape, large bin counts: because the counts are integers, each
//                           (sample, state) needs lgamma(a+j)-lgamma(a) only on an integer lattice.
//                           One CTA per SM builds the three lattices in its 200+ KB of shared memory
//                           and then streams the sample's bins through them with 128-bit loads/stores:
//                           three 8-byte shared-memory gathers + two FP64 adds per cell instead of
//                           ~150 FP64 instructions.  HBM traffic is the algorithmic 4 + 8S bytes per
//                           bin·sample (the count vector is re-read once per state, from L2).
#include "kernels.cuh"
#include "tma_ptx.cuh"

namespace edb {

// ---------------------------------------------------------------------------------------------
__global__ void state_setup_kernel(int n_samples, int n_states, const double* __restrict__ phi,
                                   const double* __restrict__ expected, const double* __restrict__ odds,
                                   StateConst* __restrict__ consts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples * n_states) return;
    const int sample = i / n_states, s = i % n_states;
    const double e = expected[sample];
    const double sd = best_sd(phi[sample], e);
    consts[i] = make_state_const(state_expected(e, odds[s]), sd);
}

void launch_state_setup(int n_samples, int n_states, const double* phi, const double* expected,
                        const double* odds, StateConst* consts, cudaStream_t st)
{
    const int n = n_samples * n_states;
    if (n == 0) return;
    state_setup_kernel<<<(n + 127) / 128, 128, 0, st>>>(n_samples, n_states, phi, expected, odds, consts);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
emission_bins_kernel(const double* __restrict__ phi, const double* __restrict__ expected,
                     const int32_t* __restrict__ total, const int32_t* __restrict__ observed, int64_t n_bins,
                     int n_states, const double* __restrict__ odds, LLView out, unsigned* __restrict__ flags, GslEventLog log)
{
    unsigned f = 0;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x) {
        const double e = expected[b];
        const double sd = best_sd(phi[b], e);
        const int tot = total[b], obs = observed[b];
        for (int s = 0; s < n_states; s++) {
            const StateConst sc = make_state_const(state_expected(e, odds[s]), sd);
            unsigned sites[2];
            out.ptr[s * out.state_stride + b] = cell_loglik_sites(sc, tot, obs, f, sites);
            if ((sites[0] | sites[1]) && log.count) {          // src/error.c:35-52: every failing call is reported
                const unsigned q = atomicAdd(log.count, 1u);
                if (q < log.cap) log.events[q] = make_uint4((unsigned)b, (unsigned)(b >> 32) << 8 | (unsigned)s, sites[0] >> kSiteShift, sites[1] >> kSiteShift);
            }
        }
    }
    if (f) atomicOr(flags, f);
}

void launch_emission_bins(const double* phi, const double* expected, const int32_t* total,
                          const int32_t* observed, int64_t n_bins, int n_states, const double* odds,
                          LLView out, unsigned* flags, GslEventLog log, cudaStream_t st)
{
    if (n_bins == 0) return;
    int64_t blocks = (n_bins + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    emission_bins_kernel<<<(int)blocks, 128, 0, st>>>(phi, expected, total, observed, n_bins, n_states, odds, out, flags, log);
}

// per-bin phi / expected for every sample of a cohort (phi.bins > 1 and covariate formulas give per-bin fits,
// R/class_definition.R:121-147, 168-180): the per-state constants are rebuilt for every bin — ~3x the arithmetic of the
// scalar kernels and 16 more bytes read per bin and sample; FP64-pipe bound
__global__ void __launch_bounds__(128)
emission_bins_batch_kernel(CountsView c, const double* __restrict__ phi, const double* __restrict__ expected, int64_t pb_stride,
                           int n_states, int64_t n_bins, const double* __restrict__ odds, LLView out, unsigned* __restrict__ flags)
{
    const int sample = blockIdx.y;
    const double* __restrict__ ph = phi + sample * pb_stride;
    const double* __restrict__ ex = expected + sample * pb_stride;
    double* o = out.ptr + sample * out.sample_stride;
    unsigned f = 0;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x) {
        int tot, obs;
        obs = c.observed[sample * c.obs_stride + b];
        const int oth = c.other[sample * c.other_stride + b];
        tot = c.other_is_total ? oth : obs + oth;
        const double e = ex[b];
        const double sd = best_sd(ph[b], e);
        for (int s = 0; s < n_states; s++) {
            const StateConst sc = make_state_const(state_expected(e, odds[s]), sd);
            o[s * out.state_stride + b] = cell_loglik(sc, tot, obs, f);
        }
    }
    if (f) atomicOr(flags, f);
}

void launch_emission_bins_batch(CountsView c, const double* phi, const double* expected, int64_t pb_stride, int n_samples,
                                int n_states, int64_t n_bins, const double* odds, LLView out, unsigned* flags, cudaStream_t st)
{
    if (n_bins == 0 || n_samples == 0) return;
    int64_t bx = (n_bins + 127) / 128;
    const int64_t want = (148 * 16 + n_samples - 1) / n_samples;
    if (bx > want) bx = want < 1 ? 1 : want;
    emission_bins_batch_kernel<<<dim3((unsigned)bx, (unsigned)n_samples), 128, 0, st>>>(c, phi, expected, pb_stride, n_states, n_bins, odds, out, flags);
}

// ---------------------------------------------------------------------------------------------
// The vendored gsl_sf_lnbeta (src/beta.c:161-164) as the device evaluates it on the faithful path: exposed so that the
// parity tests can pin the special-function chain itself (KAT-2, the dense sweeps of tests/golden/ref_vectors.npz incl.
// arguments next to negative integers: src/VP_gamma.c:795-894 through the psi / zeta closed forms).
__global__ void lnbeta_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, double* __restrict__ out,
                              unsigned* __restrict__ flags)
{
    unsigned f = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = lnbeta_gsl(x[i], y[i], f);
    if (f) atomicOr(flags, f);
}

void launch_lnbeta(const double* x, const double* y, int64_t n, double* out, unsigned* flags, cudaStream_t st)
{
    if (n == 0) return;
    int64_t blocks = (n + 127) / 128;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lnbeta_kernel<<<(int)blocks, 128, 0, st>>>(x, y, n, out, flags);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_counts(const CountsView& c, int sample, int64_t b, int& tot, int& obs)
{
    obs = c.observed[sample * c.obs_stride + b];
    const int o = c.other[sample * c.other_stride + b];
    tot = c.other_is_total ? o : obs + o;
}

__global__ void __launch_bounds__(256)
emission_direct_kernel(CountsView c, const StateConst* __restrict__ consts, int n_states, int64_t n_bins,
                       LLView out, unsigned* __restrict__ flags)
{
    __shared__ StateConst sc[kMaxStates];
    const int sample = blockIdx.y;
    for (int i = threadIdx.x; i < n_states * (int)(sizeof(StateConst) / 8); i += blockDim.x)
        reinterpret_cast<double*>(sc)[i] = reinterpret_cast<const double*>(consts + (int64_t)sample * n_states)[i];
    __syncthreads();
    unsigned f = 0;
    double* o = out.ptr + sample * out.sample_stride;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x) {
        int tot, obs;
        load_counts(c, sample, b, tot, obs);
        for (int s = 0; s < n_states; s++) o[s * out.state_stride + b] = cell_loglik(sc[s], tot, obs, f);
    }
    if (f) atomicOr(flags, f);
}

void launch_emission_direct(CountsView c, const StateConst* consts, int n_samples, int n_states,
                            int64_t n_bins, LLView out, unsigned* flags, cudaStream_t st)
{
    if (n_bins == 0 || n_samples == 0) return;
    int64_t bx = (n_bins + 255) / 256;
    const int64_t want = (148 * 8 + n_samples - 1) / n_samples;   // ~8 CTAs per SM over the whole grid
    if (bx > want) bx = want < 1 ? 1 : want;
    dim3 grid((unsigned)bx, (unsigned)n_samples);
    emission_direct_kernel<<<grid, 256, 0, st>>>(c, consts, n_states, n_bins, out, flags);
}

// ---------------------------------------------------------------------------------------------
// Table path.  One work item = (sample, state).  Lattices (see DESIGN.md "emission_table"):
//   G1[k] = lgamma(fl(a1+k))            - lgamma(a1)          k = observed
//   G2[r] = lgamma(fl(a2+r))            - lgamma(a2)          r = total - observed
//   G3[n] = lgamma(fl(a1+fl(a2+n)))     - lgamma(fl(a1+a2))   n = total
// ll = (G1[k] + G2[r]) - G3[n].  For k = 0 all three arguments coincide with the reference's own
// roundings (CNV_estimate.cpp:49); for k > 0 they can differ from them by one rounding of a sum of
// magnitude a1+a2+n, i.e. by <= ~4e-12 absolute on cells whose |ll| is then > 1.
constexpr int kTableThreads = 1024;
constexpr int kPanelEntries = 20000;   // lattices with fewer entries in total are built by the shared-index-space scheme (kPanel)
constexpr int kColdCap = 2048;      // out-of-lattice cells parked per work item before they are drained

constexpr size_t kConstSlot = (sizeof(StateConst) + 15) / 16 * 16;      // the lattices start 16-byte aligned (bulk copies)
size_t emission_table_smem_bytes(TableDims d)
{
    return sizeof(double) * ((size_t)d.K + d.R + d.N) + kConstSlot + sizeof(int) * (kColdCap + 4) + 16;
}

// Out-of-lattice cells (counts beyond the lattice, inconsistent counts) and whole pathological
// (sample, state) items are evaluated here, outside the gather loop, so that loop carries no call
// and stays within the 64 registers a 1024-thread CTA allows.
// The parking list is compacted — every lane of a draining warp has a cell — and never full: past the kColdCap entries in
// shared memory it continues in the CTA's spill area in HBM, one int per bin of the launch at most.  (Evaluating the cells
// in place, where a warp meets them, costs a full evaluation per warp and trip as soon as one lane in 32 is out of the
// lattice: counts twice as deep as the lattices were sized for took 3.95 ms instead of 0.89.)
// A parked cell usually has ONE count beyond its lattice (the test count of a deeply covered bin, or the total and the
// other count together): the terms that are in the lattice are gathered, only the others are evaluated (the lattice
// entries and gdiff are the same quantity, each to the last bit or two: DESIGN.md "emission_table").
struct LatticeView {
    const double *G1, *G2, *G3;
    TableDims dims;
};
__device__ __forceinline__ double cold_cell(const StateConst* sc, const LatticeView& lv, int tot, int obs, unsigned& f)
{
    if (obs < 0 || tot < obs) return cell_loglik(*sc, tot, obs, f);          // inconsistent counts: the reference's own NaN / error path
    double x, y, z;
    data_args(sc->a1, sc->a2, tot, obs, x, y, z);
    const int r = tot - obs;
    const double t1 = obs < lv.dims.K ? lv.G1[obs] : gdiff(sc->g1, x);
    const double t2 = r < lv.dims.R ? lv.G2[r] : gdiff(sc->g2, y);
    const double t3 = tot < lv.dims.N ? lv.G3[tot] : gdiff(sc->g12, z);
    return __dsub_rn(__dadd_rn(t1, t2), t3);
}

__device__ __noinline__ void drain_cold(const StateConst* sc, const CountsView& c, int sample, const LatticeView lv, const int* list,
                                        const int* spill, int n, double* o, unsigned* flags)
{
    unsigned f = 0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int64_t b = q < kColdCap ? list[q] : __ldcg(spill + (q - kColdCap));      // (written by this CTA: read from L2)
        int tot, obs;
        load_counts(c, sample, b, tot, obs);
        o[b] = cold_cell(sc, lv, tot, obs, f);
    }
    if (f) atomicOr(flags, f);
}

__device__ __noinline__ void whole_item_cold(const StateConst* sc, const CountsView& c, int sample, const BinRanges& rg, int64_t sub_lo,
                                             int64_t sub_hi, double* o, unsigned* flags)
{
    unsigned f = 0;
    for (int q = 0; q < rg.n; q++)
        for (int64_t b = max(rg.b0[q], sub_lo) + threadIdx.x; b < min(rg.b1[q], sub_hi); b += blockDim.x) {
            int tot, obs;
            load_counts(c, sample, b, tot, obs);
            o[b] = cell_loglik(*sc, tot, obs, f);
        }
    if (f) atomicOr(flags, f);
}

// Work items (sample, state) are dealt through an atomic counter, not by block index: when the SMs are shared with
// other kernels (the Viterbi sweep of another chromosome group) the CTAs that are resident take all the work.
// `rg`: the bin ranges this launch covers (whole matrix, or the 16-bin aligned ranges of one chromosome group).
// kPanel: small lattices for panels of a few thousand bins (many items, few bins each): the K + R + N entries are one
// index space shared evenly by ALL threads — one anchor + ~8 recurrence steps per thread — instead of one run per lattice
// and thread, whose anchors (a full lgamma difference each) would dominate a 9 K-entry build.
// kWarpRows (the full lattice kernel): a lane takes bins {2l, 2l+1} and {64+2l, 64+2l+1} of its warp's 128-bin block
// instead of four consecutive bins, so that each 128-bit store instruction of a warp covers 512 contiguous bytes.  With
// four consecutive bins per lane the two stores of an iteration each write HALF of every 32-byte sector they touch:
// ncu counted 128 M store sectors per launch where 64 M carry the data (profiles/r1i_ncu_full_summary.txt), on the L1
// data pipe that bounds the kernel; measured 0.967 -> 0.895 ms per launch (profiles/r2a_knob_ab.log).
template <bool kPanel, bool kWarpRows>
__global__ void __launch_bounds__(kTableThreads, 1)
emission_table_kernel(CountsView c, const StateConst* __restrict__ consts, int n_states, int n_items,
                      const __grid_constant__ BinRanges rg, TableDims dims, LLView out, unsigned* __restrict__ flags,
                      int* __restrict__ queue, double* __restrict__ lattices, int lattice_mode, int* __restrict__ spill_all, int64_t spill_stride,
                      int n_whole, int split)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StateConst* scp = reinterpret_cast<StateConst*>(smem_raw);
    double* G1 = reinterpret_cast<double*>(smem_raw + kConstSlot);
    double* G2 = G1 + dims.K;
    double* G3 = G2 + dims.R;
    int* cold_n = reinterpret_cast<int*>(G3 + dims.N);
    int* next_item = cold_n + 1;
    int* cold = cold_n + 4;
    int* __restrict__ spill = spill_all + (int64_t)blockIdx.x * spill_stride;
    // lattice reloads (lattice_mode 2) arrive as ONE bulk copy per item, completing on this barrier
    const uint32_t reload_bar = smem_u32(cold + kColdCap) + ((16u - (smem_u32(cold + kColdCap) & 15u)) & 15u);
    unsigned reload_phase = 0;
    if (lattice_mode == 2) {
        if (threadIdx.x == 0) {
            mbar_init(reload_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    for (;;) {
        __syncthreads();     // previous item's gathers are done before the lattices (and next_item) are overwritten
        if (threadIdx.x == 0) {
            *next_item = atomicAdd(queue, 1);
            *cold_n = 0;
        }
        __syncthreads();
        // work units: the first n_whole items whole, every later item as `split` units of consecutive bins (each builds the
        // item's lattices again): a last, partly filled round of items over the CTAs is dealt out in pieces that fill it
        const int unit = *next_item;
        if (unit >= n_whole + (n_items - n_whole) * split) break;
        int item = unit;
        int64_t sub_lo = 0, sub_hi = INT64_MAX;
        if (unit >= n_whole && split > 1) {                       // (one bin range per launch when split > 1: launch_emission_table)
            const int j = unit - n_whole, part = j % split;
            item = n_whole + j / split;
            const int64_t len = rg.b1[0] - rg.b0[0];
            sub_lo = part == 0 ? rg.b0[0] : rg.b0[0] + ((len * part / split) & ~(int64_t)127);
            sub_hi = part + 1 == split ? rg.b1[0] : rg.b0[0] + ((len * (part + 1) / split) & ~(int64_t)127);
        }
        const int sample = item / n_states, s = item - sample * n_states;
        for (int i = threadIdx.x; i < (int)(sizeof(StateConst) / 8); i += blockDim.x)
            reinterpret_cast<double*>(scp)[i] = reinterpret_cast<const double*>(consts + item)[i];
        __syncthreads();
        double* __restrict__ o = out.ptr + sample * out.sample_stride + s * out.state_stride;
        if (!scp->ok) {      // pathological shape parameters: reference NaN/sign semantics, cell by cell
            whole_item_cold(scp, c, sample, rg, sub_lo, sub_hi, o, flags);
            continue;
        }
        // lattice_mode 0: build; 1: build and keep a copy in HBM (first chromosome group of a pipelined batch);
        // 2: reload the copy (later groups: ~213 KB per item from L2 / HBM instead of ~27k lgamma differences)
        const int n_lat = dims.K + dims.R + dims.N;
        double* __restrict__ keep = lattices + (int64_t)item * n_lat;
        if (lattice_mode == 2) {
            // one bulk copy (cp.async.bulk, the TMA unit) instead of a load / store loop: that loop has one 8-byte load in
            // flight per thread, ~8 GB/s per SM against the ~1 us latency of HBM, i.e. ~26 us per 213 KB item — and a
            // pipelined batch reloads every item once per chromosome group (profiles/r2d_e2e_timeline.txt)
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy accesses to the lattices are done (barrier above)
                mbar_expect_tx(reload_bar, (unsigned)n_lat * 8u);
                tma_load_1d(smem_u32(G1), keep, (unsigned)n_lat * 8u, reload_bar);
            }
            mbar_wait(reload_bar, reload_phase);
            reload_phase ^= 1u;
        } else {
            // Every thread owns a short run of consecutive entries: the first one is a full lgamma difference, the
            // following ones use lgamma(x + 1) = lgamma(x) + log(x) — one log and a compensated addition instead of
            // ~230 instructions — as long as consecutive arguments differ by exactly 1 (they do except where a + i
            // crosses a power of two and is rounded differently; there the run is re-anchored).  The compensated sum
            // carries the run's error far below the final rounding of the entry itself.
            const GConst g1 = scp->g1, g2 = scp->g2, g12 = scp->g12;
            const double a1 = scp->a1, a2 = scp->a2;
            auto build = [&](double* __restrict__ G, int n, const GConst& g, auto arg) {
                const int per = ((n + (int)blockDim.x - 1) / (int)blockDim.x) | 1;     // odd run length: fewer bank conflicts
                int i = threadIdx.x * per;
                const int end = min(i + per, n);
                if (i >= end) return;
                double x = arg(i), s = gdiff(g, x), comp = 0.0;
                G[i] = s;
                for (++i; i < end; ++i) {
                    const double xn = arg(i);
                    if (__dadd_rn(xn, -x) == 1.0) {
                        const double t = log(x);
                        const double u = __dadd_rn(s, t), bp = __dadd_rn(u, -s);
                        comp = __dadd_rn(comp, __dadd_rn(__dadd_rn(s, -__dadd_rn(u, -bp)), __dadd_rn(t, -bp)));
                        s = u;
                    } else {
                        s = gdiff(g, xn);
                        comp = 0.0;
                    }
                    x = xn;
                    G[i] = __dadd_rn(s, comp);
                }
            };
            if constexpr (!kPanel) {
                build(G1, dims.K, g1, [&](int i) { return __dadd_rn(a1, (double)i); });
                build(G2, dims.R, g2, [&](int i) { return __dadd_rn(a2, (double)i); });
                build(G3, dims.N, g12, [&](int i) { return __dadd_rn(a1, __dadd_rn(a2, (double)i)); });
            } else {
                const int KR = dims.K + dims.R;
                const int per = (n_lat + (int)blockDim.x - 1) / (int)blockDim.x;
                int i = threadIdx.x * per;
                const int end = min(i + per, n_lat);
                double x = 0.0, s = 0.0, comp = 0.0;
                for (bool first = true; i < end; ++i, first = false) {
                    // entry i of [G1 | G2 | G3]: its lattice's constants and its argument, rounded as the gather expects
                    const GConst& g = i < dims.K ? scp->g1 : i < KR ? scp->g2 : scp->g12;
                    const double xn = i < dims.K ? __dadd_rn(a1, (double)i)
                                      : i < KR   ? __dadd_rn(a2, (double)(i - dims.K))
                                                 : __dadd_rn(a1, __dadd_rn(a2, (double)(i - KR)));
                    if (!first && i != dims.K && i != KR && __dadd_rn(xn, -x) == 1.0) {
                        const double t = log(x);
                        const double u = __dadd_rn(s, t), bp = __dadd_rn(u, -s);
                        comp = __dadd_rn(comp, __dadd_rn(__dadd_rn(s, -__dadd_rn(u, -bp)), __dadd_rn(t, -bp)));
                        s = u;
                    } else {                                    // run start, lattice boundary, or a rounding step != 1
                        s = gdiff(g, xn);
                        comp = 0.0;
                    }
                    x = xn;
                    G1[i] = __dadd_rn(s, comp);
                }
            }
        }
        __syncthreads();
        if (lattice_mode == 1)
            for (int i = threadIdx.x; i < n_lat; i += blockDim.x) keep[i] = G1[i];

        const int32_t* __restrict__ obs_row = c.observed + sample * c.obs_stride;
        const int32_t* __restrict__ oth_row = c.other + sample * c.other_stride;
        const int is_total = c.other_is_total;
        const int K = dims.K, R = dims.R, N = dims.N;

        // one cell: three gathers; out-of-lattice cells are flagged, not branched on, so that the twelve gathers
        // of a thread's four cells are in flight together (loads are clamped and unconditional)
        auto cell = [&](int obs, int oth, bool& in) -> double {
            const int tot = is_total ? oth : obs + oth;
            const int r = tot - obs;
            // unsigned compares fold the >= 0 checks in
            in = (unsigned)obs < (unsigned)K && (unsigned)r < (unsigned)R && (unsigned)tot < (unsigned)N;
            return (G1[min((unsigned)obs, (unsigned)K - 1)] + G2[min((unsigned)r, (unsigned)R - 1)]) -
                   G3[min((unsigned)tot, (unsigned)N - 1)];
        };
        auto park = [&](int64_t b) {
            const int q = atomicAdd(cold_n, 1);
            if (q < kColdCap) cold[q] = (int)b;
            else spill[q - kColdCap] = (int)b;
        };

        for (int q = 0; q < rg.n; q++) {
            // a range may start anywhere (a chromosome's first bin): scalar head up to the next multiple of 4 bins
            const int64_t rb = max(rg.b0[q], sub_lo), r1 = min(rg.b1[q], sub_hi), r0 = min(r1, (rb + 3) & ~(int64_t)3);
            if (rb >= r1) continue;
            for (int64_t bt = rb + threadIdx.x; bt < r0; bt += blockDim.x) {
                bool in;
                o[bt] = cell(obs_row[bt], oth_row[bt], in);
                if (!in) park(bt);
            }
            const bool vec = ((reinterpret_cast<uintptr_t>(obs_row + r0) | reinterpret_cast<uintptr_t>(oth_row + r0)) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(o + r0) & 15) == 0;
            if constexpr (kWarpRows) {
                const int64_t n128 = vec ? r0 + ((r1 - r0) & ~(int64_t)127) : r0;
                const int64_t stride = (int64_t)blockDim.x * 4;
                const int l2 = 2 * (int)(threadIdx.x & 31);
                int64_t b = r0 + (int64_t)(threadIdx.x >> 5) * 128 + l2;          // this lane's first pair; the second is 64 bins on
                int2 ka = make_int2(0, 0), kb = ka, oa = ka, ob = ka;
                if (b - l2 < n128) {
                    ka = __ldg(reinterpret_cast<const int2*>(obs_row + b));
                    kb = __ldg(reinterpret_cast<const int2*>(obs_row + b + 64));
                    oa = __ldg(reinterpret_cast<const int2*>(oth_row + b));
                    ob = __ldg(reinterpret_cast<const int2*>(oth_row + b + 64));
                }
                for (; b - l2 < n128; b += stride) {
                    int2 kan = make_int2(0, 0), kbn = kan, oan = kan, obn = kan;
                    if (b - l2 + stride < n128) {
                        kan = __ldg(reinterpret_cast<const int2*>(obs_row + b + stride));
                        kbn = __ldg(reinterpret_cast<const int2*>(obs_row + b + stride + 64));
                        oan = __ldg(reinterpret_cast<const int2*>(oth_row + b + stride));
                        obn = __ldg(reinterpret_cast<const int2*>(oth_row + b + stride + 64));
                    }
                    bool i0, i1, i2, i3;
                    double2 v0, v1;
                    v0.x = cell(ka.x, oa.x, i0);
                    v0.y = cell(ka.y, oa.y, i1);
                    v1.x = cell(kb.x, ob.x, i2);
                    v1.y = cell(kb.y, ob.y, i3);
                    __stcs(reinterpret_cast<double2*>(o + b), v0);
                    __stcs(reinterpret_cast<double2*>(o + b + 64), v1);
                    if (!(i0 && i1 && i2 && i3)) {                        // rare
                        if (!i0) park(b);
                        if (!i1) park(b + 1);
                        if (!i2) park(b + 64);
                        if (!i3) park(b + 65);
                    }
                    ka = kan;
                    kb = kbn;
                    oa = oan;
                    ob = obn;
                }
                for (int64_t bt = n128 + threadIdx.x; bt < r1; bt += blockDim.x) {
                    bool in;
                    o[bt] = cell(obs_row[bt], oth_row[bt], in);
                    if (!in) park(bt);
                }
                continue;
            }
            const int64_t n4 = vec ? r0 + ((r1 - r0) & ~(int64_t)3) : r0;
            // the counts of the next iteration are requested before the gathers of this one
            const int64_t stride = (int64_t)blockDim.x * 4;
            int64_t b = r0 + (int64_t)threadIdx.x * 4;
            int4 ko = make_int4(0, 0, 0, 0), oo = ko;
            if (b < n4) {
                ko = __ldg(reinterpret_cast<const int4*>(obs_row + b));
                oo = __ldg(reinterpret_cast<const int4*>(oth_row + b));
            }
            for (; b < n4; b += stride) {
                int4 kn = make_int4(0, 0, 0, 0), on = kn;
                if (b + stride < n4) {
                    kn = __ldg(reinterpret_cast<const int4*>(obs_row + b + stride));
                    on = __ldg(reinterpret_cast<const int4*>(oth_row + b + stride));
                }
                bool i0, i1, i2, i3;
                double2 v0, v1;
                v0.x = cell(ko.x, oo.x, i0);
                v0.y = cell(ko.y, oo.y, i1);
                v1.x = cell(ko.z, oo.z, i2);
                v1.y = cell(ko.w, oo.w, i3);
                __stcs(reinterpret_cast<double2*>(o + b), v0);        // streaming stores: ll is write-once
                __stcs(reinterpret_cast<double2*>(o + b + 2), v1);
                if (!(i0 && i1 && i2 && i3)) {                        // rare
                    if (!i0) park(b);
                    if (!i1) park(b + 1);
                    if (!i2) park(b + 2);
                    if (!i3) park(b + 3);
                }
                ko = kn;
                oo = on;
            }
            for (int64_t bt = n4 + threadIdx.x; bt < r1; bt += blockDim.x) {
                bool in;
                o[bt] = cell(obs_row[bt], oth_row[bt], in);
                if (!in) park(bt);
            }
        }
        __syncthreads();
        const int n_cold = *cold_n;
        const LatticeView lv{G1, G2, G3, dims};
        if (n_cold > 0) drain_cold(scp, c, sample, lv, cold, spill, n_cold, o, flags);
    }
}

int emission_table_max_ctas(int n_sms) { return 2 * n_sms; }      // panels: two 512-thread CTAs per SM

void launch_emission_table(CountsView c, const StateConst* consts, int n_samples, int n_states,
                           const BinRanges& rg, TableDims dims, LLView out, unsigned* flags, int* queue, int n_sms,
                           double* lattices, int lattice_mode, int* spill, int64_t spill_stride, cudaStream_t st)
{
    const int n_items = n_samples * n_states;
    if (n_items == 0 || rg.n == 0) return;
    cudaMemsetAsync(queue, 0, sizeof(int), st);
    const size_t smem = emission_table_smem_bytes(dims);
    if (dims.K + dims.R + dims.N < kPanelEntries) {
        static PerDevice configured;
        if (configured.raise(smem)) {
            cudaFuncSetAttribute(emission_table_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(emission_table_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        // A panel item is a chain of latency-bound phases (anchor + recurrence build, one or two gather passes, a handful
        // of out-of-lattice cells evaluated in registers): when two CTAs fit an SM's shared memory (228 KB, 1 KB reserved
        // per CTA) they run with 512 threads each — 64 registers per thread either way — and one CTA's build overlaps the
        // other's gathers.
        const bool two = 2 * (smem + 1024) <= 228 * 1024;
        const int threads = two ? kTableThreads / 2 : kTableThreads;
        const int ctas = two ? 2 * n_sms : n_sms;
        emission_table_kernel<true, false><<<n_items < ctas ? n_items : ctas, threads, smem, st>>>(c, consts, n_states, n_items, rg, dims, out, flags, queue, lattices,
                                                                                          lattices ? lattice_mode : 0, spill, spill_stride, n_items, 1);
        return;
    }
    static PerDevice configured;
    if (configured.raise(smem)) cudaFuncSetAttribute(emission_table_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // warp-row bin mapping (full-sector 128-bit stores): measured 0.967 -> 0.895 ms per launch at 256 x 200k x 5 (profiles/r2a_knob_ab.log)
    // The items left over after the whole rounds over the CTAs (64 samples x 5 states on 148 SMs: 2 rounds and 24 items) go out
    // as `split` units of consecutive bins each, every unit building the item's lattices again (~13 % of an item): the split
    // that minimises rounds x (build + gathers / split) — 24 items as 6 x 24 units cost 0.28 of a round instead of a whole one.
    int n_whole = n_items, split = 1;
    if ((!lattices || lattice_mode == 0) && rg.n == 1 && n_items % n_sms != 0) {
        const int rest = n_items % n_sms;
        const int64_t bins = rg.b1[0] - rg.b0[0];
        constexpr double kBuild = 0.13;
        double best = 1.0;
        for (int f = 2; f <= 8 && bins / f >= 16384; f++) {
            const double cost = (double)((rest * f + n_sms - 1) / n_sms) * (kBuild + (1.0 - kBuild) / f);
            if (cost < best - 0.05) {
                best = cost;
                split = f;
            }
        }
        if (split > 1) n_whole = n_items - rest;
    }
    const int n_units = n_whole + (n_items - n_whole) * split;
    emission_table_kernel<false, true><<<n_units < n_sms ? n_units : n_sms, kTableThreads, smem, st>>>(c, consts, n_states, n_items, rg, dims, out, flags, queue, lattices,
                                                                                                      lattices ? lattice_mode : 0, spill, spill_stride, n_whole, split);
}

}  // namespace edb
