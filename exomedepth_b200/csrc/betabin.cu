// betabin.cu — maximum-likelihood fit of the beta-binomial model behind new('ExomeDepth') (sm_100a, FP64).
//
// Stands in for  aod::betabin(cbind(test, reference) ~ 1, random = ~ 1, link = 'logit')  and  aod::fitted(mod)
// (R/class_definition.R:118-119, 168) for the default formula: one expected proportion mu and one over-dispersion
// phi per sample, shape parameters  a = mu (1 - phi) / phi,  b = (1 - mu)(1 - phi) / phi  (the same parametrisation
// src/CNV_estimate.cpp:45-46 uses for the normal state), maximising
//     l(a, b) = sum over bins of  lgamma(a+k) - lgamma(a) + lgamma(b+r) - lgamma(b) - [lgamma(a+b+n) - lgamma(a+b)]
// (k = test, r = reference, n = k + r; the binomial coefficient does not depend on the parameters).
// aod is a third-party package that is not part of the reference tree: PARITY UNPINNED — the definition here is "the
// maximiser of l", checked in the tests by the vanishing gradient and against an independent scipy optimiser.
//
// Because the counts are integers, lgamma(x+j) - lgamma(x) = sum_{i<j} log(x+i), so
//     l = sum_i C1[i] log(a+i) + sum_i C2[i] log(b+i) - sum_i C3[i] log(a+b+i),   C[i] = #bins whose count exceeds i
// and the gradient / Hessian are the same sums over 1/(x+i) and -1/(x+i)^2.  One CTA per sample builds the three
// "exceedance" arrays once (histogram + suffix sum in shared memory) and then runs a safeguarded Newton iteration in
// (log a, log b) whose every evaluation is a 27k-term dot product instead of a pass over 200k bins.
#include "kernels.cuh"

namespace edb {

namespace {

constexpr int kFitThreads = 1024;

// CTA-wide sums of up to 6 doubles, combined in a fixed order; every thread gets the totals
template <int NV>
__device__ void cta_sums(double (&v)[NV], double* red /* [32 * NV] */)
{
#pragma unroll
    for (int q = 0; q < NV; q++)
#pragma unroll
        for (int d = 16; d; d >>= 1) v[q] = __dadd_rn(v[q], __shfl_xor_sync(0xffffffffu, v[q], d));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; q++) red[warp * NV + q] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double s = 0.0;
        for (int w = 0; w < n_warps; w++) s = __dadd_rn(s, red[w * NV + q]);
        v[q] = s;
    }
}

// in-place inclusive suffix sum of an int array in shared memory: c[i] <- sum_{j >= i} c[j]   (n <= ~16k)
__device__ void suffix_sum(int* c, int n, int* scratch /* [blockDim.x] */)
{
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(lo + per, n);
    int s = 0;
    for (int i = hi - 1; i >= lo; i--) { s += c[i]; c[i] = s; }
    scratch[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = blockDim.x - 1; t >= 0; t--) { const int v = scratch[t]; scratch[t] = run; run += v; }
    }
    __syncthreads();
    const int add = scratch[threadIdx.x];                   // everything in the chunks after this thread's
    for (int i = lo; i < hi; i++) c[i] += add;
    __syncthreads();
}

// digamma and trigamma for x > 0: shift up to x >= 10, then the asymptotic series (error < 1e-15 relative there)
__device__ double digamma_pos(double x)
{
    double s = 0.0;
    while (x < 10.0) { s -= 1.0 / x; x += 1.0; }
    const double r = 1.0 / x, y = r * r;
    double t = 691.0 / 32760.0;
    t = fma(t, y, -1.0 / 132.0);
    t = fma(t, y, 1.0 / 240.0);
    t = fma(t, y, -1.0 / 252.0);
    t = fma(t, y, 1.0 / 120.0);
    t = fma(t, y, -1.0 / 12.0);
    return s + log(x) - 0.5 * r + t * y;
}
__device__ double trigamma_pos(double x)
{
    double s = 0.0;
    while (x < 10.0) { s += 1.0 / (x * x); x += 1.0; }
    const double r = 1.0 / x, y = r * r;
    double t = -691.0 / 2730.0;
    t = fma(t, y, 5.0 / 66.0);
    t = fma(t, y, -1.0 / 30.0);
    t = fma(t, y, 1.0 / 42.0);
    t = fma(t, y, -1.0 / 30.0);
    t = fma(t, y, 1.0 / 6.0);
    return s + r + 0.5 * y + t * y * r;
}

}  // namespace

// dims: caps of the three exceedance arrays (observed, reference, total).  Bins with a count beyond them (a handful in
// real data: ExomeCount has two) go to a per-sample overflow list and contribute through lgamma / digamma / trigamma
// differences instead; more than `ovf_cap` of them make the fit fail with info -2.
__global__ void __launch_bounds__(kFitThreads, 1)
betabin_fit_kernel(CountsView c, int n_samples, int64_t n_bins, TableDims dims, int max_iter, int2* __restrict__ overflow,
                   int ovf_cap, double* __restrict__ mu_out, double* __restrict__ phi_out, double* __restrict__ loglik_out,
                   int32_t* __restrict__ info_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* C1 = reinterpret_cast<int*>(smem_raw);
    int* C2 = C1 + dims.K;
    int* C3 = C2 + dims.R;
    int* scratch = C3 + dims.N;                              // [kFitThreads]
    double* red = reinterpret_cast<double*>(scratch + kFitThreads);     // [32 * 6]
    __shared__ int s_bad, s_novf;
    const int sample = blockIdx.x;
    int2* __restrict__ ovf = overflow + (int64_t)sample * ovf_cap;
    const int32_t* __restrict__ obs_row = c.observed + sample * c.obs_stride;
    const int32_t* __restrict__ oth_row = c.other + sample * c.other_stride;

    for (int i = threadIdx.x; i < dims.K + dims.R + dims.N; i += blockDim.x) C1[i] = 0;
    if (threadIdx.x == 0) { s_bad = 0; s_novf = 0; }
    __syncthreads();
    double sk = 0.0, sn = 0.0;
    for (int64_t b = threadIdx.x; b < n_bins; b += blockDim.x) {
        const int k = obs_row[b], o = oth_row[b];
        const int n = c.other_is_total ? o : k + o, r = n - k;
        if (k < 0 || r < 0) { s_bad = 1; continue; }
        sk += k;
        sn += n;
        if (k >= dims.K || r >= dims.R || n >= dims.N) {
            const int q = atomicAdd(&s_novf, 1);
            if (q < ovf_cap) ovf[q] = make_int2(k, r);
            continue;
        }
        // H[j] = #bins with count == j+1 contributes to C[i] for every i <= j, i.e. for i < count
        if (k > 0) atomicAdd(&C1[k - 1], 1);
        if (r > 0) atomicAdd(&C2[r - 1], 1);
        if (n > 0) atomicAdd(&C3[n - 1], 1);
    }
    __syncthreads();
    suffix_sum(C1, dims.K, scratch);
    suffix_sum(C2, dims.R, scratch);
    suffix_sum(C3, dims.N, scratch);
    double tot[2] = {sk, sn};
    cta_sums<2>(tot, red);
    const int n_ovf = s_novf;
    if (n_ovf > ovf_cap) s_bad = 1;                         // every thread writes the same value
    __syncthreads();
    if (s_bad || !(tot[1] > 0.0) || !(tot[0] > 0.0) || !(tot[0] < tot[1])) {
        if (threadIdx.x == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            mu_out[sample] = nan;
            phi_out[sample] = nan;
            loglik_out[sample] = nan;
            info_out[sample] = s_bad ? -2 : -1;             // -2: negative counts / too many bins beyond the caps; -1: degenerate sample
        }
        return;
    }
    // last non-empty entries: the dot products stop there
    int n1 = 0, n2 = 0, n3 = 0;
    {
        int m[3] = {0, 0, 0};
        for (int i = threadIdx.x; i < dims.K; i += blockDim.x) if (C1[i]) m[0] = max(m[0], i + 1);
        for (int i = threadIdx.x; i < dims.R; i += blockDim.x) if (C2[i]) m[1] = max(m[1], i + 1);
        for (int i = threadIdx.x; i < dims.N; i += blockDim.x) if (C3[i]) m[2] = max(m[2], i + 1);
        __syncthreads();
        if (threadIdx.x < 3) scratch[threadIdx.x] = 0;
        __syncthreads();
        atomicMax(&scratch[0], m[0]);
        atomicMax(&scratch[1], m[1]);
        atomicMax(&scratch[2], m[2]);
        __syncthreads();
        n1 = scratch[0];
        n2 = scratch[1];
        n3 = scratch[2];
        __syncthreads();
    }

    // value, gradient and Hessian of l in (a, b): v = {l, l_a, l_b, l_aa, l_bb, l_ab}
    auto evaluate = [&](double a, double b, double (&v)[6]) {
        const double ab = a + b;
        double l = 0, la = 0, lb = 0, laa = 0, lbb = 0, lab = 0;
        for (int i = threadIdx.x; i < n1; i += blockDim.x) {
            const double cnt = (double)C1[i], x = a + (double)i, rx = 1.0 / x;
            l = fma(cnt, log(x), l);
            la = fma(cnt, rx, la);
            laa = fma(-cnt, rx * rx, laa);
        }
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
            const double cnt = (double)C2[i], x = b + (double)i, rx = 1.0 / x;
            l = fma(cnt, log(x), l);
            lb = fma(cnt, rx, lb);
            lbb = fma(-cnt, rx * rx, lbb);
        }
        for (int i = threadIdx.x; i < n3; i += blockDim.x) {
            const double cnt = (double)C3[i], x = ab + (double)i, rx = 1.0 / x;
            l = fma(-cnt, log(x), l);
            la = fma(-cnt, rx, la);
            lb = fma(-cnt, rx, lb);
            const double q = cnt * rx * rx;
            laa += q;
            lbb += q;
            lab += q;
        }
        if (n_ovf) {                                        // the few bins beyond the arrays: closed forms
            const GConst ga = make_gconst(a), gb = make_gconst(b), gab = make_gconst(ab);
            const double pa = digamma_pos(a), pb = digamma_pos(b), pab = digamma_pos(ab);
            const double ta = trigamma_pos(a), tb = trigamma_pos(b), tab = trigamma_pos(ab);
            for (int q = threadIdx.x; q < n_ovf; q += blockDim.x) {
                const double k = (double)ovf[q].x, r = (double)ovf[q].y, n = k + r;
                l += gdiff(ga, a + k) + gdiff(gb, b + r) - gdiff(gab, ab + n);
                const double dab = digamma_pos(ab + n) - pab, tdab = trigamma_pos(ab + n) - tab;
                la += (digamma_pos(a + k) - pa) - dab;
                lb += (digamma_pos(b + r) - pb) - dab;
                laa += (trigamma_pos(a + k) - ta) - tdab;
                lbb += (trigamma_pos(b + r) - tb) - tdab;
                lab -= tdab;
            }
        }
        v[0] = l; v[1] = la; v[2] = lb; v[3] = laa; v[4] = lbb; v[5] = lab;
        cta_sums<6>(v, red);
    };

    // start: mu from the pooled proportion, phi = 0.01 (a + b = 99)
    const double mu0 = tot[0] / tot[1];
    double u = log(mu0 * 99.0), w = log((1.0 - mu0) * 99.0);          // u = log a, w = log b
    double v[6];
    evaluate(exp(u), exp(w), v);
    int it = 0, status = 1;                                 // 1: iteration cap reached
    for (; it < max_iter; it++) {
        const double a = exp(u), b = exp(w);
        // derivatives in (u, w): g = (a l_a, b l_b); H = [[a^2 l_aa + a l_a, a b l_ab], [., b^2 l_bb + b l_b]]
        const double gu = a * v[1], gw = b * v[2];
        const double huu = a * a * v[3] + gu, hww = b * b * v[4] + gw, huw = a * b * v[5];
        // converged when the gradient is negligible against the curvature scale of the problem
        if (fabs(gu) <= 1e-12 * fabs(huu) + 1e-300 && fabs(gw) <= 1e-12 * fabs(hww) + 1e-300) { status = 0; break; }
        double du, dw;
        const double det = huu * hww - huw * huw;
        if (huu < 0.0 && det > 0.0) {                       // negative definite: Newton step -H^-1 g
            du = -(hww * gu - huw * gw) / det;
            dw = -(huu * gw - huw * gu) / det;
        } else {                                            // otherwise a scaled gradient step
            const double sc = 1.0 / (fabs(huu) + fabs(hww) + 1e-300);
            du = gu * sc;
            dw = gw * sc;
        }
        const double big = fmax(fabs(du), fabs(dw));
        if (big > 2.0) { du *= 2.0 / big; dw *= 2.0 / big; }          // trust region in log space
        // backtracking on l
        double t = 1.0, vn[6];
        bool moved = false;
        for (int h = 0; h < 30; h++, t *= 0.5) {
            evaluate(exp(u + t * du), exp(w + t * dw), vn);
            if (vn[0] >= v[0] - 1e-14 * fabs(v[0])) { moved = true; break; }
        }
        if (!moved) { status = 0; break; }                  // no ascent direction left at working precision
        const double step = t * fmax(fabs(du), fabs(dw));
        u += t * du;
        w += t * dw;
#pragma unroll
        for (int q = 0; q < 6; q++) v[q] = vn[q];
        if (step < 1e-15) { status = 0; break; }
        if (u > 40.0 || w > 40.0) { status = 2; break; }    // a + b -> infinity: no over-dispersion (binomial limit)
    }
    if (threadIdx.x == 0) {
        const double a = exp(u), b = exp(w);
        mu_out[sample] = a / (a + b);
        phi_out[sample] = 1.0 / (a + b + 1.0);
        loglik_out[sample] = v[0];
        info_out[sample] = status == 0 ? it : status == 1 ? -3 : -4;   // >= 0: Newton iterations; -3: cap reached; -4: binomial limit
    }
}

// ---- expected Bayes factor of get.power.betabinom (R/tools.R:128-166; theory = FALSE, limit = FALSE) -----------------
// One CTA per problem (size, phi, p, alt.p):  sum over x = 0 .. size of  dbetabinom.ab(x | alt) * log10(e) * (log dbetabinom.ab(x | alt)
// - log dbetabinom.ab(x | null)).  The binomial coefficients cancel in the difference; every thread takes x = t, t + 256, ...
// into a compensated sum, the 256 partial sums are added in thread order (deterministic).
constexpr int kPowerThreads = 256;
__global__ void __launch_bounds__(kPowerThreads)
power_betabinom_kernel(const int32_t* __restrict__ size_in, const double* __restrict__ phi_in, const double* __restrict__ p_in,
                       const double* __restrict__ alt_in, double* __restrict__ out)
{
    __shared__ double part[kPowerThreads], comp[kPowerThreads];
    const int q = blockIdx.x;
    const int size = size_in[q];
    const double phi = phi_in[q], p = p_in[q], ap = alt_in[q];
    const double a0 = p * (1 - phi) / phi, b0 = (1 - p) * (1 - phi) / phi;          // R/tools.R:132-135
    const double a1 = ap * (1 - phi) / phi, b1 = (1 - ap) * (1 - phi) / phi;
    const double n = (double)size;
    const double k0 = lgamma(a0 + b0 + n) + (lgamma(a0) + lgamma(b0) - lgamma(a0 + b0));
    const double k1 = lgamma(a1 + b1 + n) + (lgamma(a1) + lgamma(b1) - lgamma(a1 + b1));
    const double lfn = lgamma(n + 1.0);
    double s = 0.0, c = 0.0;
    for (int x = threadIdx.x; x <= size; x += kPowerThreads) {
        const double xd = (double)x;
        const double choose = lfn - lgamma(xd + 1.0) - lgamma(n - xd + 1.0);
        const double l1 = lgamma(a1 + xd) + lgamma(b1 + n - xd) - k1, l0 = lgamma(a0 + xd) + lgamma(b0 + n - xd) - k0;
        const double t = exp(choose + l1) * (0.43429448190325182765 * (l1 - l0));
        const double u = s + t, bp = u - s;
        c += (s - (u - bp)) + (t - bp);
        s = u;
    }
    part[threadIdx.x] = s;
    comp[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, C = 0.0;
        for (int t = 0; t < kPowerThreads; t++) {
            const double v = part[t], u = S + v, bp = u - S;
            C += (S - (u - bp)) + (v - bp) + comp[t];
            S = u;
        }
        out[q] = size < 0 || !(phi > 0.0 && phi < 1.0) ? __longlong_as_double(0x7ff8000000000000LL) : S + C;
    }
}

void launch_power_betabinom(const int32_t* size, const double* phi, const double* p, const double* alt_p, int n, double* out, cudaStream_t st)
{
    if (n <= 0) return;
    prof_mark("power_betabinom", st);
    power_betabinom_kernel<<<n, kPowerThreads, 0, st>>>(size, phi, p, alt_p, out);
    prof_mark(nullptr, st);
}

size_t betabin_fit_smem_bytes(TableDims d)
{
    return sizeof(int) * ((size_t)d.K + d.R + d.N + kFitThreads) + sizeof(double) * 32 * 6 + 64;
}

void launch_betabin_fit(CountsView c, int n_samples, int64_t n_bins, TableDims dims, int max_iter, void* overflow, int ovf_cap,
                        double* mu, double* phi, double* loglik, int32_t* info, cudaStream_t st)
{
    if (n_samples == 0) return;
    const size_t smem = betabin_fit_smem_bytes(dims);
    static PerDevice configured;
    if (configured.raise(smem)) cudaFuncSetAttribute(betabin_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    prof_mark("betabin_fit", st);
    betabin_fit_kernel<<<n_samples, kFitThreads, smem, st>>>(c, n_samples, n_bins, dims, max_iter, (int2*)overflow, ovf_cap, mu, phi, loglik, info);
    prof_mark(nullptr, st);
}

}  // namespace edb
