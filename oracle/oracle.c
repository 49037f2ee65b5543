/* oracle/oracle.c — plain-C CPU restatement of ExomeDepth's emission + Viterbi hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under exomedepth_b200/ links, loads or calls this file; it is
 * the checker that tests/, __graft_entry__.smoke() and bench.py's CPU legs compare the CUDA path to.
 *
 * Pinning: every function below is checked in tests/test_oracle.py (a) bit-for-bit against the
 * reference's own C/C++ compiled unmodified (oracle/_ref/libexomedepth_ref.so) on dense sweeps, and
 * (b) against the known-answer vectors KAT-1..KAT-4 of SURVEY.md §8c committed under tests/golden/.
 * The S>3 state tables, the forward pass and the transition-probability MLE have NO counterpart in
 * the reference (hmm.cpp:37-40 rejects nstates != 3): for those this file is the definition and
 * parity is UNPINNED.
 *
 * All citations are path:line under /root/reference/.  The arithmetic below keeps the reference's
 * operation order so that, built with the same flags (-O2, no FMA contraction) against the same
 * libm, it reproduces the reference's roundings exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EDO_EPS      2.2204460492503131e-16   /* gsl_machine.h:17 */
#define EDO_ROOT4EPS 1.2207031250000000e-04   /* gsl_machine.h:20 */
#define EDO_ROOT6EPS 2.4607833005759251e-03   /* gsl_machine.h:22 */
#define EDO_PI       3.14159265358979323846264338328  /* gsl_math.h:56 */
#define EDO_E        2.71828182845904523536028747135  /* gsl_math.h:31 */
#define EDO_SQRT2    1.41421356237309504880168872421  /* gsl_math.h:43 */
#define EDO_SQRTPI   1.77245385090551602729816748334  /* gsl_math.h:68 */
#define EDO_LN2      0.69314718055994530941723212146  /* gsl_math.h:88 */
#define EDO_LNPI     1.14472988584940017414342735135  /* gsl_math.h:92 */
#define EDO_LOG_ROOT_2PI 0.9189385332046727418        /* VP_gamma.c:71 */

/* number of lnbeta evaluations that ended in a GSL-style error since the last reset (in the
 * reference each prints through src/error.c:45-48 and evaluation continues, error.c:51) */
static long g_error_events = 0;
static int g_raised = 0;
long edo_error_events(int reset)
{
    long n = g_error_events;
    if (reset) g_error_events = 0;
    return n;
}
static double domain_nan(void) { g_raised = 1; return NAN; }

/* ---- coefficient tables (values of VP_gamma.c:594-631, 637-674, 678-688; VP_log.c:73-101) ---- */
static const double GSTAR_LO[30] = {
    2.1678644786646304, -0.055332490187455841, 0.018003924314607199,
    -0.0058091926946893776, 0.0018652368948840034, -0.0005974652411395553,
    0.00019125169907783355, -6.1249965469446858e-05, 1.9638896331308425e-05,
    -6.3067741254637179e-06, 2.0288698405861392e-06, -6.5384896660838465e-07,
    2.1108698058908865e-07, -6.8260714912274945e-08, 2.2108560875880562e-08,
    -7.1710331930255456e-09, 2.3290892983985408e-09, -7.5740371598505589e-10,
    2.4658267222594333e-10, -8.0362243171659884e-11, 2.6215616826341593e-11,
    -8.5596155025948753e-12, 2.7970831499487962e-12, -9.1471771211886205e-13,
    2.9934720198063398e-13, -9.8026575909753452e-14, 3.2116773667767153e-14,
    -1.0518035333878147e-14, 3.4144405720185253e-15, -1.0115153943081187e-15,
};
static const double GSTAR_HI[30] = {
    0.0057502277273114343, 0.0004496689534965685, -0.00016727631531887174,
    6.1513701491315481e-05, -2.2372655171152501e-05, 8.0507405356647947e-06,
    -2.8671077107583396e-06, 1.0106727053742747e-06, -3.5265558477595064e-07,
    1.2179216046419402e-07, -4.1619640180795367e-08, 1.4066283500795206e-08,
    -4.6982570380537097e-09, 1.5491248664620614e-09, -5.0340936319394883e-10,
    1.6084448673736033e-10, -5.0349733196835459e-11, 1.5357154939762137e-11,
    -4.5233809655775649e-12, 1.2664429179254448e-12, -3.2648287937449326e-13,
    7.1528272726086139e-14, -9.4831735252566038e-15, -2.3124001991413208e-15,
    2.840661327717039e-15, -1.7245370321618816e-15, 8.6507923128671111e-16,
    -3.9506563665427556e-16, 1.6779342132074762e-16, -6.0483153034414767e-17,
};
static const double LANCZOS7[9] = {
    0.99999999999980993, 676.5203681218851, -1259.1392167224028,
    771.32342877765313, -176.61502916214059, 12.507343278686905,
    -0.13857109526572012, 9.9843695780195716e-06, 1.5056327351493116e-07,
};
static const double LOG1P_CHEB[21] = {
    2.1664791066439526, -0.28565398551049742, 0.015177672556905537,
    -0.0020021590494141545, 0.00019211375164056698, -2.5532588861055426e-05,
    2.9004512660400622e-06, -3.8873813517057341e-07, 4.7743678729400456e-08,
    -6.4501969776090321e-09, 8.2751976628812384e-10, -1.126049937649205e-10,
    1.4844576692270934e-11, -2.0328515972462118e-12, 2.7291231220549217e-13,
    -3.7581977830387938e-14, 5.1107345870861672e-15, -7.0722150011433277e-16,
    9.7089758328248469e-17, -1.3492637457521938e-17, 1.8657327910677295e-18,
};

/* Clenshaw sum on [-1,1] — VP_gamma.c:36-67 / VP_log.c:31-62 (error terms dropped: never consumed). */
static double clenshaw(const double *c, int order, double x)
{
    const double lo = -1.0, hi = 1.0;
    double y = (2.0 * x - lo - hi) / (hi - lo);
    double y2 = 2.0 * y;
    double d = 0.0, dd = 0.0;
    for (int j = order; j >= 1; j--) {
        double keep = d;
        d = y2 * d - dd + c[j];
        dd = keep;
    }
    return y * d - dd + 0.5 * c[0];
}

/* log(1+x) — VP_log.c:196-232 */
double edo_log1plusx(double x)
{
    if (x <= -1.0) return domain_nan();
    if (fabs(x) < EDO_ROOT6EPS) {
        const double k1 = -0.5, k2 = 1.0 / 3.0, k3 = -1.0 / 4.0, k4 = 1.0 / 5.0, k5 = -1.0 / 6.0,
                     k6 = 1.0 / 7.0, k7 = -1.0 / 8.0, k8 = 1.0 / 9.0, k9 = -1.0 / 10.0;
        double tail = k5 + x * (k6 + x * (k7 + x * (k8 + x * k9)));
        return x * (1.0 + x * (k1 + x * (k2 + x * (k3 + x * (k4 + x * tail)))));
    }
    if (fabs(x) < 0.5) {
        double t = 0.5 * (8.0 * x + 1.0) / (x + 2.0);
        return x * clenshaw(LOG1P_CHEB, 20, t);
    }
    return log(1.0 + x);
}

/* Lanczos g=7, 9 terms, real x>0 — VP_gamma.c:735-756 */
static double lanczos_lngamma(double x)
{
    x -= 1.0;
    double ag = LANCZOS7[0];
    for (int k = 1; k <= 8; k++) ag += LANCZOS7[k] / (x + k);
    double t1 = (x + 0.5) * log((x + 7.5) / EDO_E);
    double t2 = EDO_LOG_ROOT_2PI + log(ag);
    return t1 + (t2 - 7.0);
}

/* (2,2) Padé + correction for lnGamma(1+e)/e and lnGamma(2+e)/e — VP_gamma.c:928-953, 955-980 */
static double pade_near_1(double e)
{
    const double n1 = -1.0017419282349508699871138440, n2 = 1.7364839209922879823280541733;
    const double d1 = 1.2433006018858751556055436011, d2 = 5.0456274100274010152489597514;
    const double k0 = 0.004785324257581753, k1 = -0.01192457083645441, k2 = 0.01931961413960498,
                 k3 = -0.02594027398725020, k4 = 0.03141928755021455;
    double num = (e + n1) * (e + n2), den = (e + d1) * (e + d2);
    double pade = 2.0816265188662692474880210318 * num / den;
    double e5 = e * e * e * e * e;
    double corr = e5 * (k0 + e * (k1 + e * (k2 + e * (k3 + k4 * e))));
    return e * (pade + corr);
}
static double pade_near_2(double e)
{
    const double n1 = 1.000895834786669227164446568, n2 = 4.209376735287755081642901277;
    const double d1 = 2.618851904903217274682578255, d2 = 10.85766559900983515322922936;
    const double k0 = 0.0001139406357036744, k1 = -0.0001365435269792533, k2 = 0.0001067287169183665,
                 k3 = -0.0000693271800931282, k4 = 0.0000407220927867950;
    double num = (e + n1) * (e + n2), den = (e + d1) * (e + d2);
    double pade = 2.85337998765781918463568869 * num / den;
    double e5 = e * e * e * e * e;
    double corr = e5 * (k0 + e * (k1 + e * (k2 + e * (k3 + k4 * e))));
    return e * (pade + corr);
}

/* lnGamma near 0 with sign — VP_gamma.c:761-787 */
static double lngamma_near_0(double e, double *sgn)
{
    const double k1 = -0.07721566490153286061, k2 = -0.01094400467202744461,
                 k3 = 0.09252092391911371098, k4 = -0.01827191316559981266,
                 k5 = 0.01800493109685479790, k6 = -0.00685088537872380685,
                 k7 = 0.00399823955756846603, k8 = -0.00189430621687107802,
                 k9 = 0.00097473237804513221, k10 = -0.00048434392722255893;
    double g6 = k6 + e * (k7 + e * (k8 + e * (k9 + e * k10)));
    double g = e * (k1 + e * (k2 + e * (k3 + e * (k4 + e * (k5 + e * g6)))));
    double gee = g + 1.0 / (1.0 + e) + 0.5 * e;
    *sgn = e >= 0.0 ? 1.0 : -1.0;   /* GSL_SIGN, gsl_math.h */
    return log(gee / fabs(e));
}

/* Polygamma psi_n at a positive integer m (n = 0..6).  The reference reaches these through
 * gsl_sf_psi_int_e / psi_1_int_e / psi_n_e -> hzeta (VP_psi.c:604-629, 717-741, 790-816;
 * VP_zeta.c:746-806).  At integer arguments they have closed forms, restated here:
 *   psi_0(m) = -gamma + H_{m-1};  psi_n(m) = (-1)^(n+1) n! (zeta(n+1) - sum_{k<m} k^-(n+1)).
 * For large m the tail sum is replaced by its Euler–Maclaurin expansion to avoid cancellation. */
static double polygamma_int(int n, long m)
{
    static const double zeta[8] = {0, 0, 1.6449340668482264365, 1.2020569031595942854,
                                   1.0823232337111381915, 1.0369277551433699263,
                                   1.0173430619844491397, 1.0083492773819228268};
    static const double fact[7] = {1, 1, 2, 6, 24, 120, 720};
    static const double B2k[6] = {1.0 / 6, -1.0 / 30, 1.0 / 42, -1.0 / 30, 5.0 / 66, -691.0 / 2730};
    if (m > 40) {
        double x = (double)m;
        if (n == 0) {
            double s = log(x) - 0.5 / x, x2 = x * x, p = x2;
            for (int k = 1; k <= 6; k++) { s -= B2k[k - 1] / (2.0 * k * p); p *= x2; }
            return s;
        }
        /* (-1)^(n+1) [ (n-1)!/x^n + n!/(2 x^(n+1)) + sum B2k (2k+n-1)!/((2k)! x^(2k+n)) ] */
        double s = fact[n - 1] / pow(x, n) + fact[n] / (2.0 * pow(x, n + 1));
        double ratio = fact[n] * (n + 1) / 2.0; /* (2k+n-1)!/(2k)! at k=1 : (n+1)!/2 */
        for (int k = 1; k <= 6; k++) {
            s += B2k[k - 1] * ratio / pow(x, 2 * k + n);
            ratio *= (double)(2 * k + n) * (2 * k + n + 1) / ((2.0 * k + 1) * (2.0 * k + 2));
        }
        return (n & 1) ? s : -s;
    }
    if (n == 0) {
        double h = 0.0;
        for (long k = 1; k < m; k++) h += 1.0 / (double)k;
        return -0.57721566490153286061 + h;
    }
    /* sum the tail zeta(n+1, m) directly from the far end for accuracy */
    double head = 0.0;
    for (long k = m - 1; k >= 1; k--) head += pow((double)k, -(n + 1));
    double v = fact[n] * (zeta[n + 1] - head);
    return (n & 1) ? v : -v;
}

/* lnGamma for x = -N + eps, N >= 1 — VP_gamma.c:795-894 */
static double lngamma_near_negint(int N, double eps, double *sgn)
{
    if (eps == 0.0) { *sgn = 0.0; g_raised = 1; return 0.0; }
    if (N == 1) {
        const double k0 = 0.07721566490153286061, k1 = 0.08815966957356030521,
                     k2 = -0.00436125434555340577, k3 = 0.01391065882004640689,
                     k4 = -0.00409427227680839100, k5 = 0.00275661310191541584,
                     k6 = -0.00124162645565305019, k7 = 0.00065267976121802783,
                     k8 = -0.00032205261682710437, k9 = 0.00016229131039545456;
        double g5 = k5 + eps * (k6 + eps * (k7 + eps * (k8 + eps * k9)));
        double g = eps * (k0 + eps * (k1 + eps * (k2 + eps * (k3 + eps * (k4 + eps * g5)))));
        double gam_e = g - 1.0 - 0.5 * eps * (1.0 + 3.0 * eps) / (1.0 - eps * eps);
        *sgn = eps > 0.0 ? -1.0 : 1.0;
        return log(fabs(gam_e) / fabs(eps));
    }
    const double s1 = -1.6449340668482264365, s2 = 0.8117424252833536436, s3 = -0.1907518241220842137,
                 s4 = 0.0261478478176548005, s5 = -0.0023460810354558236;
    double e2 = eps * eps;
    double sin_ser = 1.0 + e2 * (s1 + e2 * (s2 + e2 * (s3 + e2 * (s4 + e2 * s5))));
    double aeps = fabs(eps);
    double c0 = lgamma((double)N + 1.0);                 /* gsl_sf_lnfact_e, VP_gamma.c:1548-1561 */
    double c1 = polygamma_int(0, (long)N + 1);
    double c2 = polygamma_int(1, (long)N + 1) / 2.0;
    double c3 = aeps > 0.00001 ? polygamma_int(2, (long)N + 1) / 6.0 : 0.0;
    double c4 = aeps > 0.0002 ? polygamma_int(3, (long)N + 1) / 24.0 : 0.0;
    double c5 = aeps > 0.001 ? polygamma_int(4, (long)N + 1) / 120.0 : 0.0;
    double c6 = aeps > 0.005 ? polygamma_int(5, (long)N + 1) / 720.0 : 0.0;
    double c7 = aeps > 0.01 ? polygamma_int(6, (long)N + 1) / 5040.0 : 0.0;
    double lng_ser = c0 - eps * (c1 - eps * (c2 - eps * (c3 - eps * (c4 - eps * (c5 - eps * (c6 - eps * c7))))));
    double g = -lng_ser - log(sin_ser);
    *sgn = ((N & 1) ? -1.0 : 1.0) * (eps > 0.0 ? 1.0 : -1.0);
    return g - log(fabs(eps));
}

/* lnGamma with sign — VP_gamma.c:1219-1285 */
double edo_lngamma_sgn(double x, double *sgn)
{
    if (fabs(x - 1.0) < 0.01) { *sgn = 1.0; return pade_near_1(x - 1.0); }
    if (fabs(x - 2.0) < 0.01) { *sgn = 1.0; return pade_near_2(x - 2.0); }
    if (x >= 0.5) { *sgn = 1.0; return lanczos_lngamma(x); }
    if (x == 0.0) { *sgn = 0.0; return domain_nan(); }
    if (fabs(x) < 0.02) return lngamma_near_0(x, sgn);
    if (x > -0.5 / (EDO_EPS * EDO_PI)) {
        double z = 1.0 - x;
        double s = sin(EDO_PI * x);
        double as = fabs(s);
        if (s == 0.0) { *sgn = 0.0; return domain_nan(); }
        if (as < EDO_PI * 0.015) {
            if (x < (double)INT32_MIN + 2.0) { *sgn = 0.0; g_raised = 1; return 0.0; }
            int N = -(int)(x - 0.5);
            double eps = x + N;
            return lngamma_near_negint(N, eps, sgn);
        }
        double lg_z = lanczos_lngamma(z);
        *sgn = s > 0.0 ? 1.0 : -1.0;
        return EDO_LNPI - (log(as) + lg_z);
    }
    *sgn = 0.0;
    g_raised = 1;
    return 0.0;
}

/* the sign-less variant used by Gamma* for x<0.5: note sin(pi*(1-x)), VP_gamma.c:1148-1216 */
static double lngamma_below_half(double x)
{
    double sgn;
    if (x == 0.0) return domain_nan();
    if (fabs(x) < 0.02) return lngamma_near_0(x, &sgn);
    double z = 1.0 - x;
    double s = sin(EDO_PI * z);
    double as = fabs(s);
    if (s == 0.0) return domain_nan();
    if (as < EDO_PI * 0.015) {
        int N = -(int)(x - 0.5);
        return lngamma_near_negint(N, x + N, &sgn);
    }
    return EDO_LNPI - (log(as) + lanczos_lngamma(z));
}

/* Temme's Gamma*(x) — VP_gamma.c:1332-1379, 986-1007 */
double edo_gammastar(double x)
{
    if (x <= 0.0) return domain_nan();
    if (x < 0.5) {
        double lg = lngamma_below_half(x);
        double lx = log(x);
        double c = 0.5 * (EDO_LN2 + EDO_LNPI);
        double lnr = lg - (x - 0.5) * lx + x - c;
        return exp(lnr);   /* gsl_sf_exp_err_e, exp.c:528-549: guards cannot fire for 0<x<0.5 */
    }
    if (x < 2.0) {
        double t = 4.0 / 3.0 * (x - 0.5) - 1.0;
        return clenshaw(GSTAR_LO, 29, t);
    }
    if (x < 10.0) {
        double t = 0.25 * (x - 2.0) - 1.0;
        double c = clenshaw(GSTAR_HI, 29, t);
        return c / (x * x) + 1.0 + 1.0 / (12.0 * x);
    }
    if (x < 1.0 / EDO_ROOT4EPS) {
        const double y = 1.0 / (x * x);
        const double q0 = 1.0 / 12.0, q1 = -1.0 / 360.0, q2 = 1.0 / 1260.0, q3 = -1.0 / 1680.0,
                     q4 = 1.0 / 1188.0, q5 = -691.0 / 360360.0, q6 = 1.0 / 156.0, q7 = -3617.0 / 122400.0;
        double ser = q0 + y * (q1 + y * (q2 + y * (q3 + y * (q4 + y * (q5 + y * (q6 + y * q7))))));
        return exp(ser / x);
    }
    if (x < 1.0 / EDO_EPS) {
        double xi = 1.0 / x;
        return 1.0 + xi / 12.0 * (1.0 + xi / 24.0 * (1.0 - xi * (139.0 / 180.0 + 571.0 / 8640.0 * xi)));
    }
    return 1.0;
}

/* ln B(x,y) — beta.c:38-47, 49-114, 161-164 */
static double lnbeta_inner(double x, double y)
{
    if (x == 0.0 || y == 0.0) return domain_nan();
    if ((x < 0 && x == floor(x)) || (y < 0 && y == floor(y))) return domain_nan();

    if (x > 0 && y > 0) {
        double mx = x > y ? x : y, mn = x < y ? x : y;
        double rat = mn / mx;
        if (rat < 0.2) {
            double gx = edo_gammastar(x), gy = edo_gammastar(y), gxy = edo_gammastar(x + y);
            double lnopr = edo_log1plusx(rat);
            double lnpre = log(gx * gy / gxy * EDO_SQRT2 * EDO_SQRTPI);
            double t1 = mn * log(rat);
            double t2 = 0.5 * log(mn);
            double t3 = (x + y - 0.5) * lnopr;
            return lnpre + (t1 - t2 - t3);
        }
    }
    double sx, sy, sxy, xy = x + y;
    double lx = edo_lngamma_sgn(x, &sx);
    double ly = edo_lngamma_sgn(y, &sy);
    double lxy = edo_lngamma_sgn(xy, &sxy);
    double val = lx + ly - lxy;
    if (sx * sy * sxy == -1.0) return domain_nan();      /* beta.c:43-45 */
    return val;
}

double edo_lnbeta(double x, double y)
{
    g_raised = 0;
    double v = lnbeta_inner(x, y);
    if (g_raised) g_error_events++;
    return v;
}

/* per-bin, per-state log-likelihood — CNV_estimate.cpp:44-50 */
static double state_loglik(double e_state, double sd, int total, int observed)
{
    double a1 = e_state * e_state * (1 - e_state) / (sd * sd) - e_state;
    double a2 = (1 - e_state) / e_state * a1;
    return edo_lnbeta(a1 + observed, a2 + total - observed) - edo_lnbeta(a1, a2);
}

/* S-state emission.  odds[s] multiplies the odds of the normal state's expected proportion;
 * reference S=3 is odds = {1-0.5*mix, 1, 1+0.5*mix} (CNV_estimate.cpp:65-66, 75-77).
 * out is column-major n x S (out[c + n*s]) like the reference's rans[] (CNV_estimate.cpp:75-77).
 * The normal state (odds == 1 exactly) passes `expected` through unchanged, as :76 does. */
void edo_emission(const double *phi, const double *expected, const int32_t *total,
                  const int32_t *observed, int64_t n, int32_t n_states, const double *odds, double *out)
{
    for (int64_t c = 0; c < n; c++) {
        double e = expected[c];
        double sd = sqrt(phi[c] * e * (1. - e));
        for (int s = 0; s < n_states; s++) {
            double es = odds[s] == 1.0 ? e : e * odds[s] / (e * odds[s] + 1 - e);
            out[c + n * s] = state_loglik(es, sd, total[c], observed[c]);
        }
    }
}

/* The reference entry point — CNV_estimate.cpp:52-85 */
void edo_get_loglike_matrix(const double *phi, const double *expected, const int32_t *total,
                            const int32_t *observed, double mixture, int64_t n, double *out)
{
    double odds[3] = {1 - 0.5 * mixture, 1.0, 1 + 0.5 * mixture};
    edo_emission(phi, expected, total, observed, n, 3, odds, out);
}

/* Distance-dependent transition row used at observation i for destination j — hmm.cpp:62-76.
 * T is the column-major S x S matrix R hands over: T[k + S*j'] ... the reference indexes
 * trans_c[j*S + k] for "from k to j". For S=3 this is exactly hmm.cpp:74-76; for S>3 rows k>=1
 * follow the same rule (extension, unpinned). */
static void transition_terms(int S, const double *T, double d, int j, double *t)
{
    t[0] = T[j * S];
    for (int k = 1; k < S; k++) t[k] = d * T[j * S + k] + (1.0 - d) * T[j * S];
}

/* log-transition table lt[(i*S + j)*S + k] = log(t_k->j at observation i), i = 1..nobs-1; row 0 unused (0).
 * Restates the host-side table of the product (exomedepth_b200/csrc/host_tables.c) for checking. */
void edo_log_transition_table(int32_t S, int32_t nobs, const double *T, const int32_t *pos, double L, double *lt)
{
    double t[16];
    for (int k = 0; k < S * S; k++) lt[k] = 0.0;
    for (int i = 1; i < nobs; i++) {
        double dist = (double)pos[i] - (double)pos[i - 1];
        double d = exp(-dist / L);
        for (int j = 0; j < S; j++) {
            transition_terms(S, T, d, j, t);
            for (int k = 0; k < S; k++) lt[((int64_t)i * S + j) * S + k] = log(t[k]);
        }
    }
}

/* Viterbi + traceback + segment summary — hmm.cpp:18-167, generalised to S states.
 * ll is column-major nobs x S in HMM state order (0 = normal).  path[nobs]; calls[4*ncalls] row-major
 * (start.p, end.p, type, nexons), 1-based like hmm.cpp:114-115.  Returns 0, or 2 when S is not in
 * [2,8].  (The reference itself refuses S != 3, hmm.cpp:37-40; callers that need that behaviour
 * check S before calling.) */
int edo_hmm(int32_t S, int32_t nobs, const double *T, const double *ll, const int32_t *pos, double L,
            int32_t *path, int32_t *calls, int32_t *ncalls)
{
    if (S < 2 || S > 8) return 2;
    double *V = (double *)malloc(sizeof(double) * 2 * S);
    int8_t *from = (int8_t *)malloc((size_t)nobs * S);
    double t[8];
    double *prev = V, *cur = V + S;
    prev[0] = 0.0;
    for (int j = 1; j < S; j++) prev[j] = -HUGE_VAL;          /* hmm.cpp:46-52 */
    for (int j = 0; j < S; j++) from[j] = -1;

    for (int i = 1; i < nobs; i++) {                           /* hmm.cpp:58 */
        double dist = (double)pos[i] - (double)pos[i - 1];
        double d = exp(-dist / L);                            /* hmm.cpp:62-64 */
        for (int j = 0; j < S; j++) {
            double best = -HUGE_VAL;
            int8_t arg = -1;
            transition_terms(S, T, d, j, t);
            double em = ll[(int64_t)j * nobs + i];
            for (int k = 0; k < S; k++) {
                double cand = em + prev[k] + log(t[k]);       /* hmm.cpp:79 */
                if (cand > best) { best = cand; arg = (int8_t)k; }   /* strict >, hmm.cpp:81 */
            }
            if (em == -HUGE_VAL) arg = 0;                     /* hmm.cpp:87 */
            cur[j] = best;
            from[(int64_t)i * S + j] = arg;
        }
        double *sw = prev; prev = cur; cur = sw;
    }

    path[nobs - 1] = 0;                                        /* hmm.cpp:95-100 */
    for (int i = nobs - 1; i >= 1; i--) {
        int st = path[i];
        /* st == -1 is the reference's latent out-of-bounds read; pinned to 0 here */
        path[i - 1] = st < 0 ? 0 : from[(int64_t)i * S + st];
    }

    int n = 0, current = 0;                                    /* hmm.cpp:104-126 */
    double start = -1., nex = 0;
    for (int i = 1; i < nobs; i++) {
        if (path[i - 1] != path[i]) {
            if (current == 0) start = i;
            if (current != 0) {
                calls[4 * n + 0] = (int32_t)(start + 1);
                calls[4 * n + 1] = i;                          /* (i-1)+1 */
                calls[4 * n + 2] = current;
                calls[4 * n + 3] = (int32_t)nex;
                n++;
                nex = 0;
            }
        }
        if (path[i] != 0) nex++;
        current = path[i];
    }
    *ncalls = n;
    free(V);
    free(from);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * EXTENSIONS with no reference counterpart (parity unpinned; this file is the definition).
 * ------------------------------------------------------------------------------------------ */

/* Forward pass in log space over the same transition model.
 * alpha_0 = (0,-Inf,..); alpha_i[j] = ll[i][j] + logsumexp_k(alpha_{i-1}[k] + log t_k->j);
 * NaN terms (log of a negative transition, SURVEY §8c "NaN edge") are skipped exactly as the
 * strict '>' of the Viterbi sweep skips them.  Returns alpha_last[0] (forced end in state 0). */
double edo_forward_loglik(int32_t S, int32_t nobs, const double *T, const double *ll, const int32_t *pos, double L)
{
    double a[8], b[8], t[8], v[8];
    a[0] = 0.0;
    for (int j = 1; j < S; j++) a[j] = -HUGE_VAL;
    for (int i = 1; i < nobs; i++) {
        double dist = (double)pos[i] - (double)pos[i - 1];
        double d = exp(-dist / L);
        for (int j = 0; j < S; j++) {
            transition_terms(S, T, d, j, t);
            double m = -HUGE_VAL;
            for (int k = 0; k < S; k++) {
                v[k] = a[k] + log(t[k]);
                if (v[k] > m) m = v[k];
            }
            if (m == -HUGE_VAL) { b[j] = -HUGE_VAL; continue; }
            double s = 0.0;
            for (int k = 0; k < S; k++) if (v[k] == v[k]) s += exp(v[k] - m);
            b[j] = ll[(int64_t)j * nobs + i] + (m + log(s));
        }
        memcpy(a, b, sizeof(double) * S);
    }
    return a[0];
}

/* CallCNVs-style S x S transition matrix for a given tp (R/class_definition.R:343-347 at S=3;
 * SURVEY §8a H4 for S>3): row 0 = (1-tp, tp/(S-1), ...), rows k>0 = 0.5 to normal, 0.5 self.
 * Column-major output, T[k + S*j] = P(k -> j). */
void edo_callcnvs_transitions(int32_t S, double tp, double *T)
{
    for (int k = 0; k < S; k++)
        for (int j = 0; j < S; j++) {
            double p;
            if (k == 0) p = j == 0 ? 1. - tp : tp / (double)(S - 1);
            else p = (j == 0 || j == k) ? 0.5 : 0.0;
            T[k + S * j] = p;
        }
}
