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

// One CTA per sample: Pearson correlation of the test and the reference counts over all bins (R/class_definition.R:338).
// The counts are integers, so the five sums n, Sx, Sy, Sxx, Syy, Sxy are accumulated EXACTLY (64-bit sums of the counts,
// 128-bit sums of their products) in ONE pass, and n*Sxy - Sx*Sy, n*Sxx - Sx^2, n*Syy - Sy^2 are exact 128-bit integers: the
// only roundings are their conversion to double, two square roots, a product and a quotient (R's two-pass long-double
// cor() agrees to the last bits).  The sums do not depend on the order of the terms.  (The two-pass compensated form this
// replaces waited 2 x 98 times on a pair of loads per thread — 0.15-0.35 ms per 64-sample chunk on SMs the emission and
// Viterbi kernels of the host pipeline were waiting for.)
struct CorSums {
    long long sx, sy;
    __int128 sxx, syy, sxy;
};
__device__ __forceinline__ double i128_to_double(__int128 v)
{
    const bool neg = v < 0;
    const unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    const unsigned long long hi = (unsigned long long)(u >> 64), lo = (unsigned long long)u;
    // hi * 2^64 is exact; the sum rounds once when hi < 2^53 (sums of products of 32-bit counts: hi < 2^40)
    const double d = fma((double)hi, 18446744073709551616.0, (double)lo);
    return neg ? -d : d;
}
__global__ void __launch_bounds__(512)
count_cor_kernel(CountsView c, int n_samples, int64_t n_bins, double* __restrict__ cor)
{
    __shared__ CorSums red[16];
    const int sample = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int32_t* __restrict__ obs = c.observed + sample * c.obs_stride;
    const int32_t* __restrict__ oth = c.other + sample * c.other_stride;
    const int is_total = c.other_is_total;
    CorSums v{0, 0, 0, 0, 0};
    auto add = [&](int o, int r) {
        const long long x = o, y = is_total ? (long long)r - o : (long long)r;
        v.sx += x;
        v.sy += y;
        v.sxx += (__int128)(x * x);          // |x| <= 2^31: x * x fits 64 bits; y = total - test can reach 2^32 in magnitude,
        v.syy += (__int128)y * y;            // so the products with y are formed in 128 bits
        v.sxy += (__int128)x * y;
    };
    // 128-bit loads, four bins per thread and trip, two trips in flight
    const bool vec = ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(oth)) & 15) == 0;
    const int64_t n4 = vec ? n_bins & ~(int64_t)3 : 0;
    const int64_t stride = (int64_t)blockDim.x * 4;
    int64_t b = (int64_t)threadIdx.x * 4;
    for (; b + stride < n4; b += 2 * stride) {
        const int4 o0 = __ldg(reinterpret_cast<const int4*>(obs + b)), r0 = __ldg(reinterpret_cast<const int4*>(oth + b));
        const int4 o1 = __ldg(reinterpret_cast<const int4*>(obs + b + stride)), r1 = __ldg(reinterpret_cast<const int4*>(oth + b + stride));
        add(o0.x, r0.x); add(o0.y, r0.y); add(o0.z, r0.z); add(o0.w, r0.w);
        add(o1.x, r1.x); add(o1.y, r1.y); add(o1.z, r1.z); add(o1.w, r1.w);
    }
    for (; b < n4; b += stride) {
        const int4 o0 = __ldg(reinterpret_cast<const int4*>(obs + b)), r0 = __ldg(reinterpret_cast<const int4*>(oth + b));
        add(o0.x, r0.x); add(o0.y, r0.y); add(o0.z, r0.z); add(o0.w, r0.w);
    }
    for (int64_t t = n4 + threadIdx.x; t < n_bins; t += blockDim.x) add(obs[t], oth[t]);
    // CTA total: integer sums, any order
    auto shfl128 = [&](__int128 x, int d) {
        const unsigned long long lo = __shfl_xor_sync(0xffffffffu, (unsigned long long)x, d);
        const unsigned long long hi = __shfl_xor_sync(0xffffffffu, (unsigned long long)((unsigned __int128)x >> 64), d);
        return (__int128)(((unsigned __int128)hi << 64) | lo);
    };
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        v.sx += __shfl_xor_sync(0xffffffffu, v.sx, d);
        v.sy += __shfl_xor_sync(0xffffffffu, v.sy, d);
        v.sxx += shfl128(v.sxx, d);
        v.syy += shfl128(v.syy, d);
        v.sxy += shfl128(v.sxy, d);
    }
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        CorSums t = red[0];
        for (int w = 1; w < n_warps; w++) {
            t.sx += red[w].sx;
            t.sy += red[w].sy;
            t.sxx += red[w].sxx;
            t.syy += red[w].syy;
            t.sxy += red[w].sxy;
        }
        const __int128 n = n_bins;
        const double num = i128_to_double(n * t.sxy - (__int128)t.sx * t.sy);
        const double dx = i128_to_double(n * t.sxx - (__int128)t.sx * t.sx), dy = i128_to_double(n * t.syy - (__int128)t.sy * t.sy);
        double r = num / (sqrt(dx) * sqrt(dy));                    // NaN for a constant vector, like R (with its warning)
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

// 12-bit layout (edb200_batch.observed12): a row is a little-endian bit stream of 12 bits per bin; a thread takes eight bins =
// 12 bytes = three 32-bit words and writes two 128-bit words.
__global__ void __launch_bounds__(256)
unpack12_counts_kernel(const uint8_t* __restrict__ src, int64_t src_stride, int32_t* __restrict__ dst, int64_t dst_stride, int64_t n_bins)
{
    const int sample = blockIdx.y;
    const uint8_t* __restrict__ s = src + sample * src_stride;
    int32_t* __restrict__ d = dst + sample * dst_stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | (uintptr_t)src_stride) & 3) == 0 && ((reinterpret_cast<uintptr_t>(d) | (uintptr_t)(dst_stride * 4)) & 15) == 0;
    const int64_t n8 = vec ? n_bins / 8 : 0;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n8; g += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s + 12 * g);
        const uint32_t w0 = __ldcs(w), w1 = __ldcs(w + 1), w2 = __ldcs(w + 2);
        reinterpret_cast<int4*>(d + 8 * g)[0] = make_int4((int)(w0 & 0xFFFu), (int)((w0 >> 12) & 0xFFFu), (int)((w0 >> 24) | ((w1 & 0xFu) << 8)), (int)((w1 >> 4) & 0xFFFu));
        reinterpret_cast<int4*>(d + 8 * g)[1] = make_int4((int)((w1 >> 16) & 0xFFFu), (int)((w1 >> 28) | ((w2 & 0xFFu) << 4)), (int)((w2 >> 8) & 0xFFFu), (int)(w2 >> 20));
    }
    for (int64_t b = n8 * 8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t at = (3 * b) >> 1;                                   // byte of bit 12 b
        d[b] = (b & 1) ? (int)(s[at] >> 4) | ((int)s[at + 1] << 4) : (int)s[at] | (((int)s[at + 1] & 0xF) << 8);
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

int launch_unpack12_counts(const uint8_t* src, int64_t src_stride, int32_t* dst, int64_t dst_stride, int n_samples, int64_t n_bins, cudaStream_t st)
{
    if (n_samples == 0 || n_bins == 0) return 0;
    int bx = (int)((n_bins / 8 + 255) / 256);
    bx = bx < 1 ? 1 : bx > 64 ? 64 : bx;
    prof_mark("widen_counts", st);
    unpack12_counts_kernel<<<dim3((unsigned)bx, (unsigned)n_samples), 256, 0, st>>>(src, src_stride, dst, dst_stride, n_bins);
    prof_mark(nullptr, st);
    return 1;
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
