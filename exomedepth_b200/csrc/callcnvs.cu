// callcnvs.cu — the numeric columns CallCNVs adds to every call, computed where the data already is (sm_100a).
//
// Replaces the per-call R loop of R/class_definition.R:393-403 and the correlation of :338:
//   BF              sum over the call's bins of  ll[, type] - ll[, normal]          (:395-397; x log10(e) and signif on the host)
//   reads.expected  sum over the call's bins of  total * expected                   (:399; as.integer on the host)
//   reads.observed  sum over the call's bins of  test                               (:400)
//   cor(test, reference) over all bins of the sample                                (:338)
// so that the FP64 likelihood matrix (8*S bytes per bin and sample) does not have to cross PCIe just to be summed
// over a few hundred short segments.  R's sum() accumulates in long double; here every sum is a compensated
// (two-sum) FP64 accumulation, which carries more than the 64 mantissa bits of x87 long double, combined in a fixed
// order: the results are deterministic and agree with an exactly rounded sum except in the last bit of rare cases.
#include <cstdint>
#include "kernels.cuh"

namespace edb {

namespace {

struct Comp {                      // running compensated sum: value = s + c
    double s, c, plain;            // `plain` is the naive sum, returned when s + c is not finite (Inf / NaN terms)
};
__device__ __forceinline__ void comp_add(Comp& a, double x)
{
    const double t = __dadd_rn(a.s, x);
    const double bp = __dadd_rn(t, -a.s);
    const double err = __dadd_rn(__dadd_rn(a.s, -__dadd_rn(t, -bp)), __dadd_rn(x, -bp));
    a.s = t;
    a.c = __dadd_rn(a.c, err);
    a.plain = __dadd_rn(a.plain, x);
}
__device__ __forceinline__ void comp_merge(Comp& a, double s2, double c2, double p2)
{
    const double t = __dadd_rn(a.s, s2);
    const double bp = __dadd_rn(t, -a.s);
    const double err = __dadd_rn(__dadd_rn(a.s, -__dadd_rn(t, -bp)), __dadd_rn(s2, -bp));
    a.s = t;
    a.c = __dadd_rn(__dadd_rn(a.c, c2), err);
    a.plain = __dadd_rn(a.plain, p2);
}
__device__ __forceinline__ double comp_warp_total(Comp a)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double s2 = __shfl_xor_sync(0xffffffffu, a.s, d), c2 = __shfl_xor_sync(0xffffffffu, a.c, d);
        const double p2 = __shfl_xor_sync(0xffffffffu, a.plain, d);
        // both partners must merge in the same order to hold the same value: lower lane's sum first
        Comp lo = a, hi{s2, c2, p2};
        if (threadIdx.x & d) { lo = hi; hi = a; }
        comp_merge(lo, hi.s, hi.c, hi.plain);
        a = lo;
    }
    const double v = __dadd_rn(a.s, a.c);
    return (v - v == 0.0) ? v : a.plain;
}

}  // namespace

// One warp per (sample, call).  calls: (start.p, end.p, type, nexons) with 1-based GLOBAL bin indices (CallCNVs
// numbering after its -1 and per-chromosome shift); a call whose start.p is < 1 sums nothing (oracle/framing.py).
__global__ void __launch_bounds__(128)
call_summary_kernel(CallSummaryArgs a)
{
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int sample = (int)(wid / a.call_cap), k = (int)(wid - (int64_t)sample * a.call_cap);
    if (sample >= a.n_samples) return;
    const int n = a.ncalls[sample];
    if (k >= n) return;
    const int32_t* __restrict__ call = a.calls + ((int64_t)sample * a.call_cap + k) * 4;
    const int sp = call[0], ep = call[1], typ = call[2];
    double* __restrict__ out = a.stats + ((int64_t)sample * a.call_cap + k) * 3;
    Comp bf{0, 0, 0}, rexp{0, 0, 0}, robs{0, 0, 0};
    if (sp >= 1 && ep <= a.n_bins && typ >= 0 && typ < a.n_states) {
        const double* __restrict__ llt = a.ll + sample * a.ll_sample_stride + (int64_t)a.perm[typ] * a.ll_state_stride;
        const double* __restrict__ lln = a.ll + sample * a.ll_sample_stride + (int64_t)a.perm[0] * a.ll_state_stride;
        const int32_t* __restrict__ obs = a.counts.observed + sample * a.counts.obs_stride;
        const int32_t* __restrict__ oth = a.counts.other + sample * a.counts.other_stride;
        // per-sample scalar, or one value per bin (edb200_batch.per_bin_stride: R/class_definition.R:168-180)
        const double* __restrict__ ev = a.expected + (int64_t)sample * (a.expected_stride ? a.expected_stride : 1);
        for (int64_t b = sp - 1 + lane; b < ep; b += 32) {
            const double e = a.expected_stride ? ev[b] : ev[0];
            comp_add(bf, __dadd_rn(llt[b], -lln[b]));
            const int o = obs[b];
            const double tot = a.counts.other_is_total ? (double)oth[b] : __dadd_rn((double)o, (double)oth[b]);
            comp_add(rexp, __dmul_rn(tot, e));
            comp_add(robs, (double)o);
        }
    }
    const double v0 = comp_warp_total(bf), v1 = comp_warp_total(rexp), v2 = comp_warp_total(robs);
    if (lane == 0) {
        out[0] = v0;
        out[1] = v1;
        out[2] = v2;
    }
}

// One CTA per sample: Pearson correlation of the test and the reference counts over all bins, two passes like R's
// cor() (means first, then centred sums), every sum compensated.
__global__ void __launch_bounds__(512)
count_cor_kernel(CountsView c, int n_samples, int64_t n_bins, double* __restrict__ cor)
{
    __shared__ double red[16][3][3];
    __shared__ double mean[2];
    const int sample = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int32_t* __restrict__ obs = c.observed + sample * c.obs_stride;
    const int32_t* __restrict__ oth = c.other + sample * c.other_stride;
    // the CTA-wide total of up to three compensated sums, combined in warp order
    auto cta_total = [&](Comp* v, int nv, double* out) {
        for (int q = 0; q < nv; q++) {
            Comp t = v[q];
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double s2 = __shfl_xor_sync(0xffffffffu, t.s, d), c2 = __shfl_xor_sync(0xffffffffu, t.c, d);
                const double p2 = __shfl_xor_sync(0xffffffffu, t.plain, d);
                Comp lo = t, hi{s2, c2, p2};
                if (threadIdx.x & d) { lo = hi; hi = t; }
                comp_merge(lo, hi.s, hi.c, hi.plain);
                t = lo;
            }
            if (lane == 0) { red[warp][q][0] = t.s; red[warp][q][1] = t.c; red[warp][q][2] = t.plain; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < nv; q++) {
                Comp t{0, 0, 0};
                for (int w = 0; w < n_warps; w++) comp_merge(t, red[w][q][0], red[w][q][1], red[w][q][2]);
                out[q] = __dadd_rn(t.s, t.c);
            }
        }
        __syncthreads();
    };
    // 128-bit loads, four bins per thread and trip (with one 4-byte load per trip the kernel is bound by the latency of
    // 2 x 391 dependent trips: 0.2 ms for 64 samples x 200k bins, during which its CTAs keep the emission kernel of the next
    // chunk off their SMs)
    const bool vec = ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(oth)) & 15) == 0;
    const int64_t n4 = vec ? n_bins & ~(int64_t)3 : 0;
    auto each = [&](auto&& f) {
        for (int64_t b = (int64_t)threadIdx.x * 4; b < n4; b += (int64_t)blockDim.x * 4) {
            const int4 o = __ldg(reinterpret_cast<const int4*>(obs + b)), r = __ldg(reinterpret_cast<const int4*>(oth + b));
            f(o.x, r.x);
            f(o.y, r.y);
            f(o.z, r.z);
            f(o.w, r.w);
        }
        for (int64_t b = n4 + threadIdx.x; b < n_bins; b += blockDim.x) f(obs[b], oth[b]);
    };
    const int is_total = c.other_is_total;
    Comp v[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    each([&](int o, int r) {
        comp_add(v[0], (double)o);
        comp_add(v[1], is_total ? __dadd_rn((double)r, -(double)o) : (double)r);
    });
    __shared__ double tot[3];
    cta_total(v, 2, tot);
    if (threadIdx.x == 0) {
        mean[0] = tot[0] / (double)n_bins;
        mean[1] = tot[1] / (double)n_bins;
    }
    __syncthreads();
    const double mx = mean[0], my = mean[1];
    v[0] = v[1] = v[2] = Comp{0, 0, 0};
    each([&](int o, int r) {
        const double dx = __dadd_rn((double)o, -mx), dy = __dadd_rn(is_total ? __dadd_rn((double)r, -(double)o) : (double)r, -my);
        comp_add(v[0], __dmul_rn(dx, dy));
        comp_add(v[1], __dmul_rn(dx, dx));
        comp_add(v[2], __dmul_rn(dy, dy));
    });
    cta_total(v, 3, tot);
    if (threadIdx.x == 0) {
        double r = tot[0] / (sqrt(tot[1]) * sqrt(tot[2]));       // NaN for a constant vector, like R (with its warning)
        if (r > 1.0) r = 1.0;
        if (r < -1.0) r = -1.0;
        cor[sample] = n_bins > 1 ? r : __longlong_as_double(0x7ff8000000000000ll);
    }
}

int launch_call_summary(const CallSummaryArgs& a, cudaStream_t st)
{
    int launches = 0;
    if (a.n_samples == 0) return 0;
    if (a.stats) {
        const int64_t warps = (int64_t)a.n_samples * a.call_cap;
        prof_mark("call_summary", st);
        call_summary_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(a);
        launches++;
    }
    if (a.cor) {
        prof_mark("count_cor", st);
        count_cor_kernel<<<a.n_samples, 512, 0, st>>>(a.counts, a.n_samples, a.n_bins, a.cor);
        launches++;
    }
    prof_mark(nullptr, st);
    return launches;
}

// ---- count ingestion: 16-bit counts + overflow list -> the int32 rows the kernels read ---------------
// Exome read counts per bin fit 16 bits except for a handful of bins per sample: the ingestion layout is uint16
// [sample][bin] with 65535 standing for "see the overflow list" (sorted flat indices sample * n_bins + bin, int32 values).
// It halves the bytes that cross PCIe per sample; the device widens the columns of one chromosome group right behind
// their upload, on the copy stream.
__global__ void __launch_bounds__(256)
widen_counts_kernel(const uint16_t* __restrict__ src, int64_t src_stride, int32_t* __restrict__ dst, int64_t dst_stride, int n_samples,
                    const __grid_constant__ BinRanges rg)
{
    const int sample = blockIdx.y;
    const uint16_t* __restrict__ s = src + sample * src_stride;
    int32_t* __restrict__ d = dst + sample * dst_stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | (uintptr_t)(src_stride * 2)) & 15) == 0 &&
                     ((reinterpret_cast<uintptr_t>(d) | (uintptr_t)(dst_stride * 4)) & 15) == 0;
    for (int q = 0; q < rg.n; q++) {
        const int64_t b1 = rg.b1[q], b0 = min(b1, (rg.b0[q] + 7) & ~(int64_t)7);      // scalar head up to a multiple of 8 bins
        for (int64_t b = rg.b0[q] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < b0; b += (int64_t)gridDim.x * blockDim.x) d[b] = s[b];
        const int64_t n8 = vec ? b0 + ((b1 - b0) & ~(int64_t)7) : b0;
        for (int64_t b = b0 + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; b < n8; b += (int64_t)gridDim.x * blockDim.x * 8) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(s + b));
            reinterpret_cast<int4*>(d + b)[0] = make_int4((int)(v.x & 0xFFFFu), (int)(v.x >> 16), (int)(v.y & 0xFFFFu), (int)(v.y >> 16));
            reinterpret_cast<int4*>(d + b)[1] = make_int4((int)(v.z & 0xFFFFu), (int)(v.z >> 16), (int)(v.w & 0xFFFFu), (int)(v.w >> 16));
        }
        for (int64_t b = n8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < b1; b += (int64_t)gridDim.x * blockDim.x) d[b] = s[b];
    }
}

__global__ void patch_overflow_kernel(const int64_t* __restrict__ index, const int32_t* __restrict__ value, int64_t n_overflow, int64_t n_bins,
                                      int32_t* __restrict__ dst, int64_t dst_stride, const __grid_constant__ BinRanges rg, int64_t s_lo, int64_t s_hi)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_overflow) return;
    const int64_t sample = index[i] / n_bins, b = index[i] - sample * n_bins;
    if (sample < s_lo || sample >= s_hi) return;
    for (int q = 0; q < rg.n; q++)
        if (b >= rg.b0[q] && b < rg.b1[q]) dst[sample * dst_stride + b] = value[i];
}

int launch_widen_counts(const uint16_t* src, int64_t src_stride, int32_t* dst, int64_t dst_stride, int n_samples, int64_t n_bins,
                        const BinRanges& rg, const int64_t* ovf_index, const int32_t* ovf_value, int64_t n_overflow, cudaStream_t st)
{
    if (n_samples == 0 || rg.n == 0) return 0;
    int64_t width = 0;
    for (int q = 0; q < rg.n; q++) width += rg.b1[q] - rg.b0[q];
    int bx = (int)((width / 8 + 255) / 256);
    bx = bx < 1 ? 1 : bx > 64 ? 64 : bx;
    prof_mark("widen_counts", st);
    widen_counts_kernel<<<dim3((unsigned)bx, (unsigned)n_samples), 256, 0, st>>>(src, src_stride, dst, dst_stride, n_samples, rg);
    int launches = 1;
    if (n_overflow > 0) {
        // (dst is row 0 of the batch here: the list's flat indices are absolute)
        patch_overflow_kernel<<<(unsigned)((n_overflow + 127) / 128), 128, 0, st>>>(ovf_index, ovf_value, n_overflow, n_bins, dst, dst_stride, rg, 0, INT64_MAX);
        launches++;
    }
    prof_mark(nullptr, st);
    return launches;
}

// the overflow entries alone, over the bins of `rg` and the samples s_lo .. s_hi-1 (dst = row 0 of the batch)
int launch_patch_overflow(int32_t* dst, int64_t dst_stride, int64_t n_bins, const BinRanges& rg, const int64_t* ovf_index,
                          const int32_t* ovf_value, int64_t n_overflow, cudaStream_t st, int64_t s_lo, int64_t s_hi)
{
    if (n_overflow <= 0) return 0;
    patch_overflow_kernel<<<(unsigned)((n_overflow + 127) / 128), 128, 0, st>>>(ovf_index, ovf_value, n_overflow, n_bins, dst, dst_stride, rg, s_lo, s_hi);
    return 1;
}

}  // namespace edb
