// refset.cu — the correlation sweep of select.reference.set (sm_100a, FP64).
//
// Replaces  my.correlations <- apply(reference.counts, 2, function(x) cor(x/(bin.length*sum(x)/10^6),
//                                                     test.counts/(bin.length*sum(test.counts)/10^6)))
// (R/optimize_reference_set.R:100) for every sample of a cohort at once: with leave-one-out cohorts the selected bins
// do not depend on which sample is the test, so the sweep is one Pearson matrix of the normalised count rows
// (SURVEY.md §8f-1).  Two kernels:
//   refset_standardize_kernel  one thread-block cluster of 8 CTAs per sample (sums exchanged through distributed shared
//                              memory): gather the selected bins, y = x / (bin.length * sum(x) / 1e6)
//                              evaluated as R does, two-pass mean / centred norm with compensated sums,
//                              z = (y - mean) / norm  ->  Z[sample][bin], K padded with zeros.  HBM-bound (reads
//                              the int32 counts once per pass, writes 8 bytes per selected bin).
//   refset_gram_kernel         C = Za . Zb^T (rows of Za against rows of Zb over the selected bins), FP64 FMA on a
//                              128 x 128 x 16 shared-memory tiling, 8 x 8 outputs per thread, split over K so that a
//                              256-sample cohort still fills the 148 SMs; the K-slices are summed in slice order by
//                              refset_reduce_kernel (deterministic, no atomics).  This is the one dense contraction of
//                              the package; it stays on the FP64 pipe because the result is compared at 1e-10.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace edb {

namespace {

struct Acc {                       // compensated running sum (two-sum), value = s + c
    double s, c;
};
__device__ __forceinline__ void acc_add(Acc& a, double x)
{
    const double t = __dadd_rn(a.s, x);
    const double bp = __dadd_rn(t, -a.s);
    a.c = __dadd_rn(a.c, __dadd_rn(__dadd_rn(a.s, -__dadd_rn(t, -bp)), __dadd_rn(x, -bp)));
    a.s = t;
}
}  // namespace

// One thread-block CLUSTER of kStdCluster CTAs per sample: every CTA takes one contiguous slice of the selected bins; the
// three sample-wide sums (total, mean, centred norm) are exchanged through distributed shared memory — each CTA
// publishes its compensated partial sum in its own shared memory, the cluster synchronises, and every CTA adds the
// partials of all ranks in rank order, so all of them hold the same bits.
constexpr int kStdCluster = 8, kStdThreads = 512;

__device__ double cluster_total(Acc a, double* red /* [64] */, double* slot /* [2], this CTA's published partial */)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    // CTA-wide partial, compensated, in lane / warp order
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double s2 = __shfl_xor_sync(0xffffffffu, a.s, d), c2 = __shfl_xor_sync(0xffffffffu, a.c, d);
        Acc lo = a, hi{s2, c2};
        if (threadIdx.x & d) { lo = hi; hi = a; }
        acc_add(lo, hi.s);
        lo.c = __dadd_rn(lo.c, hi.c);
        a = lo;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    if (lane == 0) { red[2 * warp] = a.s; red[2 * warp + 1] = a.c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Acc t{0, 0};
        for (int w = 0; w < n_warps; w++) {
            acc_add(t, red[2 * w]);
            t.c = __dadd_rn(t.c, red[2 * w + 1]);
        }
        slot[0] = t.s;
        slot[1] = t.c;
    }
    cluster.sync();                                         // every CTA's partial is published
    Acc t{0, 0};
    for (unsigned r = 0; r < cluster.num_blocks(); r++) {
        const double* remote = cluster.map_shared_rank(slot, r);
        acc_add(t, remote[0]);
        t.c = __dadd_rn(t.c, remote[1]);
    }
    cluster.sync();                                         // nobody overwrites its slot while others still read it
    return __dadd_rn(t.s, t.c);
}

__global__ void __cluster_dims__(kStdCluster, 1, 1) __launch_bounds__(kStdThreads)
refset_standardize_kernel(const int32_t* __restrict__ counts, int64_t stride, const double* __restrict__ bin_length,
                          const int32_t* __restrict__ selected, int64_t n_sel, int64_t k_pad, double* __restrict__ z)
{
    __shared__ double red[64];
    __shared__ double slot[2];
    const int sample = blockIdx.x / kStdCluster, part = blockIdx.x % kStdCluster;
    const int32_t* __restrict__ row = counts + sample * stride;
    double* __restrict__ out = z + sample * k_pad;          // also the scratch for y between the passes (L2 resident)
    const int64_t per = (k_pad / 16 + kStdCluster - 1) / kStdCluster * 16;      // slice of this CTA, whole 16-bin groups
    const int64_t i0 = part * per, i1 = i0 + per < k_pad ? i0 + per : k_pad, s1 = i1 < n_sel ? i1 : n_sel;
    Acc a{0, 0};
    for (int64_t i = i0 + threadIdx.x; i < s1; i += blockDim.x) acc_add(a, (double)row[selected[i]]);
    const double total = cluster_total(a, red, slot);       // sum(x) over the selected bins: an integer, exact
    a = Acc{0, 0};
    for (int64_t i = i0 + threadIdx.x; i < s1; i += blockDim.x) {
        const int32_t b = selected[i];
        const double bl = bin_length ? bin_length[b] : 1.0;
        const double y = __ddiv_rn((double)row[b], __ddiv_rn(__dmul_rn(bl, total), 1e6));   // x / ((bin.length * sum(x)) / 10^6)
        out[i] = y;
        acc_add(a, y);
    }
    const double mean = __ddiv_rn(cluster_total(a, red, slot), (double)n_sel);
    a = Acc{0, 0};
    for (int64_t i = i0 + threadIdx.x; i < s1; i += blockDim.x) {      // each thread re-reads what it wrote itself
        const double d = __dadd_rn(out[i], -mean);
        out[i] = d;
        acc_add(a, __dmul_rn(d, d));
    }
    const double norm = sqrt(cluster_total(a, red, slot));  // 0 for a constant row: z becomes NaN, cor() gives NA there too
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) out[i] = i < n_sel ? __ddiv_rn(out[i], norm) : 0.0;
}

// The B operand may live on several GPUs: row j of the cohort is row j % rows_per_rank of rank j / rows_per_rank, and
// `zb.base[rank]` is that rank's standardised block mapped into this process (CUDA IPC; loads go over NVLink / NVSwitch
// peer memory).  The all-gather of the sharded sweep is thereby fused into the contraction: tiles are fetched from
// their owners while other tiles are being multiplied, and no rank ever holds a copy of the whole matrix Z.
constexpr int kGT = 128, kGK = 16;
constexpr int kGS = kGT + 4;           // shared-memory row stride in doubles: the four k-rows of a fragment load fall on distinct banks
// FP64 tensor-core contraction (mma.sync m8n8k4, DMMA): a warp forms a 64 x 32 block of the 128 x 128 tile as 8 x 4
// fragments of 8 x 8; per 4 values of k it loads 8 + 4 operand fragments (one double per lane each) for 32 DMMAs, where
// the FMA form read 16 doubles per lane for 64 FMAs and stalled on its register operands (19 of 34 TFLOP/s, with or
// without double buffering).  Two shared-memory stages: the next K-chunk's global loads (this rank's rows, and the peers'
// over NVLink) are issued before the current chunk's DMMAs and stashed into the other stage behind them: one barrier per
// chunk.  The sum over k runs in the tensor unit's order inside a group of 4 and in ascending groups — fixed, so the
// sharded forms of the sweep still agree bit for bit.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
refset_gram_kernel(const double* __restrict__ za, int m, const __grid_constant__ PeerRows zb, int n, int64_t k_pad, int64_t k_slice,
                   double* __restrict__ partial)
{
    extern __shared__ __align__(16) double gram_smem[];     // [stage][a | b][kGK][kGS], k-major
    auto sa = [&](int st, int k, int r) -> double& { return gram_smem[((st * 2 + 0) * kGK + k) * kGS + r]; };
    auto sb = [&](int st, int k, int r) -> double& { return gram_smem[((st * 2 + 1) * kGK + k) * kGS + r]; };
    const int ti = blockIdx.y * kGT, tj = blockIdx.x * kGT;
    const int64_t k0 = (int64_t)blockIdx.z * k_slice, k1 = k0 + k_slice < k_pad ? k0 + k_slice : k_pad;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = (warp >> 2) * 64, wx = (warp & 3) * 32;             // this warp's block of the tile
    const int fr = lane >> 2, fk = lane & 3;                           // fragment row / k of this lane
    // loads: thread t brings 8 consecutive k of one row of each operand tile (128 rows x 16 k = 256 threads x 8)
    const int lr = threadIdx.x >> 1, lk = (threadIdx.x & 1) * 8;
    const double2* pa = ti + lr < m ? reinterpret_cast<const double2*>(za + (int64_t)(ti + lr) * k_pad + lk) : nullptr;
    const double2* pb = nullptr;
    if (tj + lr < n) {
        const int row = tj + lr, owner = row / zb.rows_per_rank;
        pb = reinterpret_cast<const double2*>(zb.base[owner] + (int64_t)(row - owner * zb.rows_per_rank) * k_pad + lk);
    }
    double2 av[4], bv[4];
    auto fetch = [&](int64_t k) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            av[q] = pa ? __ldg(pa + k / 2 + q) : make_double2(0, 0);
            bv[q] = pb ? pb[k / 2 + q] : make_double2(0, 0);
        }
    };
    auto stash = [&](int st) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            sa(st, lk + 2 * q, lr) = av[q].x; sa(st, lk + 2 * q + 1, lr) = av[q].y;
            sb(st, lk + 2 * q, lr) = bv[q].x; sb(st, lk + 2 * q + 1, lr) = bv[q].y;
        }
    };
    double acc[8][4][2] = {};
    int cur = 0;
    if (k0 < k1) {
        fetch(k0);
        stash(0);
    }
    __syncthreads();
    for (int64_t k = k0; k < k1; k += kGK) {
        const bool more = k + kGK < k1;
        if (more) fetch(k + kGK);
#pragma unroll
        for (int k4 = 0; k4 < kGK; k4 += 4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = sa(cur, k4 + fk, wy + i * 8 + fr);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sb(cur, k4 + fk, wx + j * 8 + fr);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (more) stash(cur ^ 1);       // (the other stage: nobody reads it during this chunk)
        __syncthreads();
        cur ^= 1;
    }
    double* __restrict__ out = partial + (int64_t)blockIdx.z * m * n;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int r = ti + wy + i * 8 + fr, c = tj + wx + j * 8 + 2 * fk + e;
                if (r < m && c < n) out[(int64_t)r * n + c] = acc[i][j][e];
            }
}

__global__ void __launch_bounds__(256)
refset_reduce_kernel(const double* __restrict__ partial, int n_slices, int64_t mn, double* __restrict__ c)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= mn) return;
    double s = 0.0;
    for (int q = 0; q < n_slices; q++) s = __dadd_rn(s, partial[(int64_t)q * mn + i]);
    c[i] = s > 1.0 ? 1.0 : s < -1.0 ? -1.0 : s;            // R's cor() clamps to [-1, 1]; NaN passes through
}

void launch_refset_standardize(const int32_t* counts, int64_t stride, int n_samples, const double* bin_length,
                               const int32_t* selected, int64_t n_sel, int64_t k_pad, double* z, cudaStream_t st)
{
    if (n_samples == 0) return;
    prof_mark("refset_standardize", st);
    refset_standardize_kernel<<<n_samples * kStdCluster, kStdThreads, 0, st>>>(counts, stride, bin_length, selected, n_sel, k_pad, z);
    prof_mark(nullptr, st);
}

// K-slices of the Gram kernel.  The slice length depends on K only — not on the block of rows a GPU forms — so that an
// entry of the matrix is the same sum in the same order however the samples are sharded: at most 64 slices of at
// least 1024 bins.
int refset_gram_slices(int m, int n, int64_t k_pad, int n_sms)
{
    (void)m; (void)n; (void)n_sms;
    int64_t k_slice = (k_pad + 63) / 64;
    if (k_slice < 1024) k_slice = 1024;
    k_slice = (k_slice + kGK - 1) / kGK * kGK;
    return (int)((k_pad + k_slice - 1) / k_slice);
}

void launch_refset_gram(const double* za, int m, const PeerRows& zb, int n, int64_t k_pad, int n_slices, double* partial,
                        double* c, cudaStream_t st)
{
    if (m == 0 || n == 0) return;
    int64_t k_slice = (k_pad + 63) / 64;                    // as in refset_gram_slices
    if (k_slice < 1024) k_slice = 1024;
    k_slice = (k_slice + kGK - 1) / kGK * kGK;
    prof_mark("refset_gram", st);
    constexpr size_t kGramSmem = 2 * 2 * kGK * kGS * sizeof(double);      // 66 KB
    static PerDevice configured;
    if (configured.raise(kGramSmem)) cudaFuncSetAttribute(refset_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem);
    refset_gram_kernel<<<dim3((n + kGT - 1) / kGT, (m + kGT - 1) / kGT, n_slices), 256, kGramSmem, st>>>(za, m, zb, n, k_pad, k_slice, partial);
    prof_mark("refset_reduce", st);
    const int64_t mn = (int64_t)m * n;
    refset_reduce_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, st>>>(partial, n_slices, mn, c);
    prof_mark(nullptr, st);
}

}  // namespace edb
