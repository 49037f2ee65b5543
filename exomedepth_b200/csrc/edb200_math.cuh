// edb200_math.cuh — FP64 device math for the beta-binomial emission log-likelihood (sm_100a).
//
// Two evaluators of   ll = lnB(a1+k, a2+n-k) - lnB(a1, a2)   (reference: src/CNV_estimate.cpp:44-50):
//
//  * the PRODUCT path (`gdiff_*`): for a1,a2 > 0 the six lgamma terms are regrouped into three
//    differences  lgamma(x) - lgamma(a)  that are evaluated without cancellation
//        (x-1/2) log1p((x-a)/a) + (x-a)(ln a - 1) + [S(x) - S(a)],   S = Stirling tail,
//    so the absolute error scales with the read counts instead of with a1+a2 (which reaches several
//    thousand for small phi and is what limits the reference's own accuracy, SURVEY.md §7 hard part 3).
//    Everything that depends only on (phi, expected, state) is hoisted into a StateConst.
//
//  * the FAITHFUL path (`lnbeta_gsl`): a branch-for-branch device restatement of the vendored GSL
//    routines (src/beta.c:49-114, src/VP_gamma.c:735-756, 761-787, 795-894, 928-980, 1219-1285,
//    1332-1379, src/VP_log.c:196-232).  It is only taken when a shape parameter is not positive
//    (pathological phi) or the counts are inconsistent, where the reference's
//    NaN / sign behaviour has to be reproduced cell by cell.
//
// No tensor cores: scalar special functions (BASELINE.json north_star).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace edb {

constexpr int    kMaxStates = 7;
constexpr double kStirlingMin = 10.0;   // 8-term Stirling tail is exact to < 2e-18 for x >= 10

// status bits (mirrors include/exomedepth_b200.h)
constexpr unsigned kFlagNaN = 1u;       // a GSL-style domain error produced NaN cells (reference prints + continues)
// ... and WHICH error site of the vendored GSL chain raised it (src/error.c:35-52 prints file, line and reason of every
// failing call and carries on): one bit per site, in the order the reference reaches them inside one gsl_sf_lnbeta call.
// capi.cu: kGslSites holds the file / line / reason of every bit; edb200_gsl_error_log prints them like the reference.
constexpr unsigned kSiteShift = 8;
constexpr unsigned kSiteBetaZero = 1u << 8;         // beta.c:56      x == 0 || y == 0
constexpr unsigned kSiteBetaNegInt = 1u << 9;       // beta.c:59      negative integer argument
constexpr unsigned kSiteGammastar = 1u << 10;       // VP_gamma.c:1338  gammastar(x <= 0)
constexpr unsigned kSiteLog1p = 1u << 11;           // VP_log.c:202   log_1plusx(x <= -1)
constexpr unsigned kSiteLngammaZero = 1u << 12;     // VP_gamma.c:1239  lngamma_sgn(0)
constexpr unsigned kSiteLngammaSin = 1u << 13;      // VP_gamma.c:1253  sin(pi x) == 0
constexpr unsigned kSiteLngammaRound1 = 1u << 14;   // VP_gamma.c:1261  "error" (x < INT_MIN + 2)
constexpr unsigned kSiteNearNegInt = 1u << 15;      // VP_gamma.c:803   "error" (eps == 0)
constexpr unsigned kSiteLngammaRound2 = 1u << 16;   // VP_gamma.c:1283  "error" (|x| too large)
constexpr unsigned kSiteBetaSign = 1u << 17;        // beta.c:44      the beta function is negative
constexpr unsigned kSiteMask = 0x3FFu << 8;

// ------------------------------------------------------------------------------------------------
// product path
// ------------------------------------------------------------------------------------------------

// log(1+u) for u > -1 with RELATIVE accuracy for small |u| (u = (x-a)/a is known to ~2 ulp).
__device__ __forceinline__ double log1p_rel(double u)
{
    // explicit roundings: every instance of this code (lattice build, in-register kernel, cold path) must produce
    // the same bits, so nothing here is left to the compiler's mul+add contraction
    const double w = __dadd_rn(1.0, u);
    const double c = __dsub_rn(u, __dsub_rn(w, 1.0));   // rounding error of 1+u (exact for |u| < 2^52)
    return __dadd_rn(log(w), __ddiv_rn(c, w));
}

// Stirling tail S(x) = lgamma(x) - [(x-1/2)ln x - x + ln(2pi)/2], x >= kStirlingMin
__device__ __forceinline__ double stirling_tail(double x)
{
    const double r = __ddiv_rn(1.0, x);
    const double y = __dmul_rn(r, r);
    double s = -3617.0 / 122400.0;
    s = fma(s, y, 1.0 / 156.0);
    s = fma(s, y, -691.0 / 360360.0);
    s = fma(s, y, 1.0 / 1188.0);
    s = fma(s, y, -1.0 / 1680.0);
    s = fma(s, y, 1.0 / 1260.0);
    s = fma(s, y, -1.0 / 360.0);
    s = fma(s, y, 1.0 / 12.0);
    return __dmul_rn(s, r);
}

// lgamma(x) for any finite x > 0: shift up to the Stirling range.
__device__ __forceinline__ double lgamma_pos(double x)
{
    double p = 1.0;
    while (x < kStirlingMin) { p = __dmul_rn(p, x); x = __dadd_rn(x, 1.0); }
    const double v = __dadd_rn(__dadd_rn(fma(__dsub_rn(x, 0.5), log(x), -x), 0.91893853320467274178), stirling_tail(x));
    return p == 1.0 ? v : __dsub_rn(v, log(p));
}

// Hoisted constants of one shape parameter a > 0.
struct GConst {
    double a;       // the shape parameter
    double ra;      // 1/a
    double lnam1;   // ln(a) - 1
    double tail;    // S(a)            (a >= kStirlingMin)
    double lg;      // lgamma(a)       (a <  kStirlingMin)
    int    small;   // a < kStirlingMin
};

__device__ __forceinline__ GConst make_gconst(double a)
{
    GConst c;
    c.a = a;
    c.ra = __ddiv_rn(1.0, a);
    c.small = !(a >= kStirlingMin);
    c.lnam1 = __dsub_rn(log(a), 1.0);
    c.tail = c.small ? 0.0 : stirling_tail(a);
    c.lg = c.small ? lgamma_pos(a) : 0.0;
    return c;
}

// lgamma(x) - lgamma(a) for x > 0 (x is a plus a non-negative integer up to rounding).
__device__ __forceinline__ double gdiff(const GConst& c, double x)
{
    const double d = __dsub_rn(x, c.a);
    if (d == 0.0) return 0.0;
    if (!c.small && x >= kStirlingMin) {
        const double L = log1p_rel(__dmul_rn(d, c.ra));
        return fma(__dsub_rn(x, 0.5), L, fma(d, c.lnam1, __dsub_rn(stirling_tail(x), c.tail)));
    }
    const double lga = c.small ? c.lg : lgamma_pos(c.a);
    return __dsub_rn(lgamma_pos(x), lga);
}

// Per (sample, state) constants.
struct StateConst {
    GConst g1, g2, g12;   // a1, a2, fl(a1+a2)
    int    ok;            // 1: a1, a2 finite and > 0 -> product path; 0 -> faithful path
    double a1, a2;        // as the reference computes them
};

// Shape parameters exactly as src/CNV_estimate.cpp:45-46 computes them (no FMA contraction).
__device__ __forceinline__ void shape_params(double e_state, double sd, double& a1, double& a2)
{
    const double one_m = __dsub_rn(1.0, e_state);
    const double num = __dmul_rn(__dmul_rn(e_state, e_state), one_m);
    a1 = __dsub_rn(__ddiv_rn(num, __dmul_rn(sd, sd)), e_state);
    a2 = __dmul_rn(__ddiv_rn(one_m, e_state), a1);
}

// Expected proportion of a state with the given odds (src/CNV_estimate.cpp:75,77); odds == 1 passes
// `e` through unchanged like :76 does.
__device__ __forceinline__ double state_expected(double e, double odds)
{
    if (odds == 1.0) return e;
    const double eo = __dmul_rn(e, odds);
    return __ddiv_rn(eo, __dsub_rn(__dadd_rn(eo, 1.0), e));
}

__device__ __forceinline__ double best_sd(double phi, double e)
{
    return __dsqrt_rn(__dmul_rn(__dmul_rn(phi, e), __dsub_rn(1.0, e)));   // CNV_estimate.cpp:73
}

__device__ __forceinline__ StateConst make_state_const(double e_state, double sd)
{
    StateConst s;
    shape_params(e_state, sd, s.a1, s.a2);
    s.ok = (s.a1 > 0.0) && (s.a2 > 0.0) && (s.a1 < 1e300) && (s.a2 < 1e300);
    if (s.ok) {
        s.g1 = make_gconst(s.a1);
        s.g2 = make_gconst(s.a2);
        s.g12 = make_gconst(__dadd_rn(s.a1, s.a2));
    }
    return s;
}

// The three arguments of the data term with the reference's own roundings
// (`a1 + observed`, `a2 + total - observed`, and x+y inside lnbeta; CNV_estimate.cpp:49, beta.c:102).
__device__ __forceinline__ void data_args(double a1, double a2, int total, int observed,
                                          double& x, double& y, double& z)
{
    x = __dadd_rn(a1, (double)observed);
    y = __dsub_rn(__dadd_rn(a2, (double)total), (double)observed);
    z = __dadd_rn(x, y);
}

// ------------------------------------------------------------------------------------------------
// faithful path: device restatement of the vendored GSL subset
// ------------------------------------------------------------------------------------------------
namespace gsl {

constexpr double kEps = 2.2204460492503131e-16;
constexpr double kPi = 3.14159265358979323846264338328;
constexpr double kLnPi = 1.14472988584940017414342735135;

static __device__ const double GSTAR_LO[30] = {
    2.1678644786646304, -0.055332490187455841, 0.018003924314607199,
    -0.0058091926946893776, 0.0018652368948840034, -0.0005974652411395553,
    0.00019125169907783355, -6.1249965469446858e-05, 1.9638896331308425e-05,
    -6.3067741254637179e-06, 2.0288698405861392e-06, -6.5384896660838465e-07,
    2.1108698058908865e-07, -6.8260714912274945e-08, 2.2108560875880562e-08,
    -7.1710331930255456e-09, 2.3290892983985408e-09, -7.5740371598505589e-10,
    2.4658267222594333e-10, -8.0362243171659884e-11, 2.6215616826341593e-11,
    -8.5596155025948753e-12, 2.7970831499487962e-12, -9.1471771211886205e-13,
    2.9934720198063398e-13, -9.8026575909753452e-14, 3.2116773667767153e-14,
    -1.0518035333878147e-14, 3.4144405720185253e-15, -1.0115153943081187e-15};
static __device__ const double GSTAR_HI[30] = {
    0.0057502277273114343, 0.0004496689534965685, -0.00016727631531887174,
    6.1513701491315481e-05, -2.2372655171152501e-05, 8.0507405356647947e-06,
    -2.8671077107583396e-06, 1.0106727053742747e-06, -3.5265558477595064e-07,
    1.2179216046419402e-07, -4.1619640180795367e-08, 1.4066283500795206e-08,
    -4.6982570380537097e-09, 1.5491248664620614e-09, -5.0340936319394883e-10,
    1.6084448673736033e-10, -5.0349733196835459e-11, 1.5357154939762137e-11,
    -4.5233809655775649e-12, 1.2664429179254448e-12, -3.2648287937449326e-13,
    7.1528272726086139e-14, -9.4831735252566038e-15, -2.3124001991413208e-15,
    2.840661327717039e-15, -1.7245370321618816e-15, 8.6507923128671111e-16,
    -3.9506563665427556e-16, 1.6779342132074762e-16, -6.0483153034414767e-17};
static __device__ const double LANCZOS7[9] = {
    0.99999999999980993, 676.5203681218851, -1259.1392167224028, 771.32342877765313,
    -176.61502916214059, 12.507343278686905, -0.13857109526572012, 9.9843695780195716e-06,
    1.5056327351493116e-07};
static __device__ const double LOG1P_CHEB[21] = {
    2.1664791066439526, -0.28565398551049742, 0.015177672556905537,
    -0.0020021590494141545, 0.00019211375164056698, -2.5532588861055426e-05,
    2.9004512660400622e-06, -3.8873813517057341e-07, 4.7743678729400456e-08,
    -6.4501969776090321e-09, 8.2751976628812384e-10, -1.126049937649205e-10,
    1.4844576692270934e-11, -2.0328515972462118e-12, 2.7291231220549217e-13,
    -3.7581977830387938e-14, 5.1107345870861672e-15, -7.0722150011433277e-16,
    9.7089758328248469e-17, -1.3492637457521938e-17, 1.8657327910677295e-18};

// Clenshaw recurrence on [-1,1] (VP_gamma.c:36-67)
static __device__ __noinline__ double clenshaw(const double* c, int order, double x)
{
    const double y = x, y2 = 2.0 * x;
    double d = 0.0, dd = 0.0;
    for (int j = order; j >= 1; j--) {
        const double keep = d;
        d = y2 * d - dd + c[j];
        dd = keep;
    }
    return y * d - dd + 0.5 * c[0];
}

// VP_log.c:196-232
static __device__ __noinline__ double log_1plusx(double x, unsigned& flags)
{
    if (x <= -1.0) { flags |= kFlagNaN | kSiteLog1p; return nan(""); }
    if (fabs(x) < 2.4607833005759251e-03) {
        double t = -1.0 / 6.0 + x * (1.0 / 7.0 + x * (-1.0 / 8.0 + x * (1.0 / 9.0 + x * (-1.0 / 10.0))));
        return x * (1.0 + x * (-0.5 + x * (1.0 / 3.0 + x * (-0.25 + x * (0.2 + x * t)))));
    }
    if (fabs(x) < 0.5) {
        const double t = 0.5 * (8.0 * x + 1.0) / (x + 2.0);
        return x * clenshaw(LOG1P_CHEB, 20, t);
    }
    return log(1.0 + x);
}

// VP_gamma.c:735-756
static __device__ __noinline__ double lanczos_lngamma(double x)
{
    x -= 1.0;
    double ag = LANCZOS7[0];
    for (int k = 1; k <= 8; k++) ag += LANCZOS7[k] / (x + k);
    const double t1 = (x + 0.5) * log((x + 7.5) / 2.71828182845904523536028747135);
    const double t2 = 0.9189385332046727418 + log(ag);
    return t1 + (t2 - 7.0);
}

// VP_gamma.c:928-953 and 955-980
__device__ __forceinline__ double pade_near(double e, double n1, double n2, double d1, double d2, double scale,
                                            double k0, double k1, double k2, double k3, double k4)
{
    const double num = (e + n1) * (e + n2), den = (e + d1) * (e + d2);
    const double pade = scale * num / den;
    const double e5 = e * e * e * e * e;
    const double corr = e5 * (k0 + e * (k1 + e * (k2 + e * (k3 + k4 * e))));
    return e * (pade + corr);
}

// VP_gamma.c:761-787
__device__ __forceinline__ double lngamma_near_0(double e, double& sgn)
{
    const double g6 = -0.00685088537872380685 +
                      e * (0.00399823955756846603 +
                           e * (-0.00189430621687107802 + e * (0.00097473237804513221 + e * -0.00048434392722255893)));
    const double g = e * (-0.07721566490153286061 +
                          e * (-0.01094400467202744461 +
                               e * (0.09252092391911371098 +
                                    e * (-0.01827191316559981266 + e * (0.01800493109685479790 + e * g6)))));
    const double gee = g + 1.0 / (1.0 + e) + 0.5 * e;
    sgn = e >= 0.0 ? 1.0 : -1.0;
    return log(gee / fabs(e));
}

// psi_n at a positive integer m (closed forms; the reference reaches these through
// VP_psi.c:604-629, 717-741, 790-816 and VP_zeta.c:746-806). Same restatement as oracle/oracle.c.
static __device__ __noinline__ double polygamma_int(int n, long long m)
{
    const double zeta[8] = {0, 0, 1.6449340668482264365, 1.2020569031595942854, 1.0823232337111381915,
                            1.0369277551433699263, 1.0173430619844491397, 1.0083492773819228268};
    const double fact[7] = {1, 1, 2, 6, 24, 120, 720};
    const double B2k[6] = {1.0 / 6, -1.0 / 30, 1.0 / 42, -1.0 / 30, 5.0 / 66, -691.0 / 2730};
    if (m > 40) {
        const double x = (double)m;
        if (n == 0) {
            double s = log(x) - 0.5 / x, x2 = x * x, p = x2;
            for (int k = 1; k <= 6; k++) { s -= B2k[k - 1] / (2.0 * k * p); p *= x2; }
            return s;
        }
        double s = fact[n - 1] / pow(x, (double)n) + fact[n] / (2.0 * pow(x, (double)(n + 1)));
        double ratio = fact[n] * (n + 1) / 2.0;
        for (int k = 1; k <= 6; k++) {
            s += B2k[k - 1] * ratio / pow(x, (double)(2 * k + n));
            ratio *= (double)(2 * k + n) * (2 * k + n + 1) / ((2.0 * k + 1) * (2.0 * k + 2));
        }
        return (n & 1) ? s : -s;
    }
    if (n == 0) {
        double h = 0.0;
        for (long long k = 1; k < m; k++) h += 1.0 / (double)k;
        return -0.57721566490153286061 + h;
    }
    double head = 0.0;
    for (long long k = m - 1; k >= 1; k--) head += pow((double)k, -(double)(n + 1));
    const double v = fact[n] * (zeta[n + 1] - head);
    return (n & 1) ? v : -v;
}

// VP_gamma.c:795-894 (x = -N + eps)
static __device__ __noinline__ double lngamma_near_negint(int N, double eps, double& sgn, unsigned& flags)
{
    if (eps == 0.0) { sgn = 0.0; flags |= kFlagNaN | kSiteNearNegInt; return 0.0; }
    if (N == 1) {
        const double g5 = 0.00275661310191541584 +
                          eps * (-0.00124162645565305019 +
                                 eps * (0.00065267976121802783 + eps * (-0.00032205261682710437 + eps * 0.00016229131039545456)));
        const double g = eps * (0.07721566490153286061 +
                                eps * (0.08815966957356030521 +
                                       eps * (-0.00436125434555340577 +
                                              eps * (0.01391065882004640689 + eps * (-0.00409427227680839100 + eps * g5)))));
        const double gam_e = g - 1.0 - 0.5 * eps * (1.0 + 3.0 * eps) / (1.0 - eps * eps);
        sgn = eps > 0.0 ? -1.0 : 1.0;
        return log(fabs(gam_e) / fabs(eps));
    }
    const double e2 = eps * eps;
    const double sin_ser = 1.0 + e2 * (-1.6449340668482264365 +
                                       e2 * (0.8117424252833536436 +
                                             e2 * (-0.1907518241220842137 + e2 * (0.0261478478176548005 + e2 * -0.0023460810354558236))));
    const double aeps = fabs(eps);
    const long long m = (long long)N + 1;
    const double c0 = lgamma_pos((double)N + 1.0);
    const double c1 = polygamma_int(0, m);
    const double c2 = polygamma_int(1, m) / 2.0;
    const double c3 = aeps > 0.00001 ? polygamma_int(2, m) / 6.0 : 0.0;
    const double c4 = aeps > 0.0002 ? polygamma_int(3, m) / 24.0 : 0.0;
    const double c5 = aeps > 0.001 ? polygamma_int(4, m) / 120.0 : 0.0;
    const double c6 = aeps > 0.005 ? polygamma_int(5, m) / 720.0 : 0.0;
    const double c7 = aeps > 0.01 ? polygamma_int(6, m) / 5040.0 : 0.0;
    const double ser = c0 - eps * (c1 - eps * (c2 - eps * (c3 - eps * (c4 - eps * (c5 - eps * (c6 - eps * c7))))));
    const double g = -ser - log(sin_ser);
    sgn = ((N & 1) ? -1.0 : 1.0) * (eps > 0.0 ? 1.0 : -1.0);
    return g - log(fabs(eps));
}

// VP_gamma.c:1219-1285.  use_1mx selects sin(pi*(1-x)) as the sign-less variant (:1148-1216) does.
static __device__ __noinline__ double lngamma_sgn(double x, double& sgn, unsigned& flags, bool use_1mx = false)
{
    if (fabs(x - 1.0) < 0.01) {
        sgn = 1.0;
        return pade_near(x - 1.0, -1.0017419282349508699871138440, 1.7364839209922879823280541733,
                         1.2433006018858751556055436011, 5.0456274100274010152489597514,
                         2.0816265188662692474880210318, 0.004785324257581753, -0.01192457083645441,
                         0.01931961413960498, -0.02594027398725020, 0.03141928755021455);
    }
    if (fabs(x - 2.0) < 0.01) {
        sgn = 1.0;
        return pade_near(x - 2.0, 1.000895834786669227164446568, 4.209376735287755081642901277,
                         2.618851904903217274682578255, 10.85766559900983515322922936,
                         2.85337998765781918463568869, 0.0001139406357036744, -0.0001365435269792533,
                         0.0001067287169183665, -0.0000693271800931282, 0.0000407220927867950);
    }
    if (x >= 0.5) { sgn = 1.0; return lanczos_lngamma(x); }
    if (x == 0.0) { sgn = 0.0; flags |= kFlagNaN | kSiteLngammaZero; return nan(""); }
    if (fabs(x) < 0.02) return lngamma_near_0(x, sgn);
    if (x > -0.5 / (kEps * kPi)) {
        const double z = 1.0 - x;
        const double s = sin(kPi * (use_1mx ? z : x));
        const double as = fabs(s);
        if (s == 0.0) { sgn = 0.0; flags |= kFlagNaN | kSiteLngammaSin; return nan(""); }
        if (as < kPi * 0.015) {
            if (x < -2147483646.0) { sgn = 0.0; flags |= kFlagNaN | kSiteLngammaRound1; return 0.0; }
            const int N = -(int)(x - 0.5);
            return lngamma_near_negint(N, x + N, sgn, flags);
        }
        sgn = s > 0.0 ? 1.0 : -1.0;
        return kLnPi - (log(as) + lanczos_lngamma(z));
    }
    sgn = 0.0;
    flags |= kFlagNaN | kSiteLngammaRound2;
    return 0.0;
}

// VP_gamma.c:1332-1379, 986-1007
static __device__ __noinline__ double gammastar(double x, unsigned& flags)
{
    if (x <= 0.0) { flags |= kFlagNaN | kSiteGammastar; return nan(""); }
    if (x < 0.5) {
        double sgn;
        const double lg = lngamma_sgn(x, sgn, flags, true);
        const double lx = log(x);
        const double c = 0.5 * (0.69314718055994530941723212146 + kLnPi);
        return exp(lg - (x - 0.5) * lx + x - c);
    }
    if (x < 2.0) return clenshaw(GSTAR_LO, 29, 4.0 / 3.0 * (x - 0.5) - 1.0);
    if (x < 10.0) {
        const double c = clenshaw(GSTAR_HI, 29, 0.25 * (x - 2.0) - 1.0);
        return c / (x * x) + 1.0 + 1.0 / (12.0 * x);
    }
    if (x < 1.0 / 1.2207031250000000e-04) {
        const double y = 1.0 / (x * x);
        const double ser = 1.0 / 12.0 +
                           y * (-1.0 / 360.0 +
                                y * (1.0 / 1260.0 +
                                     y * (-1.0 / 1680.0 +
                                          y * (1.0 / 1188.0 + y * (-691.0 / 360360.0 + y * (1.0 / 156.0 + y * (-3617.0 / 122400.0)))))));
        return exp(ser / x);
    }
    if (x < 1.0 / kEps) {
        const double xi = 1.0 / x;
        return 1.0 + xi / 12.0 * (1.0 + xi / 24.0 * (1.0 - xi * (139.0 / 180.0 + 571.0 / 8640.0 * xi)));
    }
    return 1.0;
}

}  // namespace gsl

// beta.c:38-47, 49-114, 161-164
static __device__ __noinline__ double lnbeta_gsl(double x, double y, unsigned& flags)
{
    if (x == 0.0 || y == 0.0) { flags |= kFlagNaN | kSiteBetaZero; return nan(""); }
    if ((x < 0 && x == floor(x)) || (y < 0 && y == floor(y))) { flags |= kFlagNaN | kSiteBetaNegInt; return nan(""); }
    if (x > 0 && y > 0) {
        const double mx = fmax(x, y), mn = fmin(x, y);
        const double rat = mn / mx;
        if (rat < 0.2) {
            const double gx = gsl::gammastar(x, flags), gy = gsl::gammastar(y, flags), gxy = gsl::gammastar(x + y, flags);
            const double lnopr = gsl::log_1plusx(rat, flags);
            const double lnpre = log(gx * gy / gxy * 1.41421356237309504880168872421 * 1.77245385090551602729816748334);
            const double t1 = mn * log(rat);
            const double t2 = 0.5 * log(mn);
            const double t3 = (x + y - 0.5) * lnopr;
            return lnpre + (t1 - t2 - t3);
        }
    }
    double sx, sy, sxy;
    const double lx = gsl::lngamma_sgn(x, sx, flags);
    const double ly = gsl::lngamma_sgn(y, sy, flags);
    const double lxy = gsl::lngamma_sgn(x + y, sxy, flags);
    if (sx * sy * sxy == -1.0) { flags |= kFlagNaN | kSiteBetaSign; return nan(""); }
    return lx + ly - lxy;
}

// ------------------------------------------------------------------------------------------------
// one cell of the emission matrix
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double cell_loglik(const StateConst& sc, int total, int observed, unsigned& flags)
{
    double x, y, z;
    data_args(sc.a1, sc.a2, total, observed, x, y, z);
    if (sc.ok && observed >= 0 && total >= observed) {
        // (G1 + G2) - G3 ; identical grouping in the table path so both agree bit for bit when x,y,z agree
        return __dsub_rn(__dadd_rn(gdiff(sc.g1, x), gdiff(sc.g2, y)), gdiff(sc.g12, z));
    }
    return lnbeta_gsl(x, y, flags) - lnbeta_gsl(sc.a1, sc.a2, flags);
}

// the same cell with the error sites of its two gsl_sf_lnbeta calls reported apart (sites[0]: the data term, sites[1]: the
// normalising term) — the `.Call`-shaped entry point logs them per cell like the reference prints them
__device__ __forceinline__ double cell_loglik_sites(const StateConst& sc, int total, int observed, unsigned& flags, unsigned* sites)
{
    double x, y, z;
    data_args(sc.a1, sc.a2, total, observed, x, y, z);
    sites[0] = sites[1] = 0u;
    if (sc.ok && observed >= 0 && total >= observed)
        return __dsub_rn(__dadd_rn(gdiff(sc.g1, x), gdiff(sc.g2, y)), gdiff(sc.g12, z));
    const double t = lnbeta_gsl(x, y, sites[0]), n = lnbeta_gsl(sc.a1, sc.a2, sites[1]);
    flags |= (sites[0] | sites[1]) & kFlagNaN;
    return t - n;
}

}  // namespace edb
