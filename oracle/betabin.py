"""scipy restatement of the beta-binomial fit of new('ExomeDepth').  TEST INFRASTRUCTURE ONLY.

The reference fits with aod::betabin(cbind(test, reference) ~ 1, random = ~ 1, link = 'logit')
(R/class_definition.R:118-119) — a third-party package without a version pin (DESCRIPTION:9) whose sources are not
part of the reference tree: PARITY UNPINNED.  What is checked instead is the definition both share: (mu, phi) maximise
    sum over bins of  lbeta(a + test, b + reference) - lbeta(a, b),   a = mu (1-phi)/phi,  b = (1-mu)(1-phi)/phi
(aod's parametrisation; the binomial coefficients are constant).  SURVEY.md §8c records an independent scipy fit of
ExomeCount (Exome4 against Exome1+2+3): expected ~ 0.2176, phi ~ 0.00451."""
import numpy as np
from scipy import optimize, special


def loglik(a, b, test, reference):
    test = np.asarray(test, float)
    reference = np.asarray(reference, float)
    return float(np.sum(special.betaln(a + test, b + reference) - special.betaln(a, b)))


def gradient_log(a, b, test, reference):
    """d loglik / d(log a), d loglik / d(log b)."""
    test = np.asarray(test, float)
    reference = np.asarray(reference, float)
    n = test + reference
    common = np.sum(special.digamma(a + b + n) - special.digamma(a + b))
    ga = np.sum(special.digamma(a + test) - special.digamma(a)) - common
    gb = np.sum(special.digamma(b + reference) - special.digamma(b)) - common
    return a * ga, b * gb


def fit(test, reference):
    """(mu, phi, loglik) by an independent optimiser: BFGS with the analytic (digamma) gradient in (log a, log b) on the
    likelihood scaled to O(1)."""
    test = np.asarray(test, float)
    reference = np.asarray(reference, float)
    mu0 = test.sum() / (test.sum() + reference.sum())
    x0 = np.log([mu0 * 99.0, (1 - mu0) * 99.0])

    def f(x):
        return -loglik(np.exp(x[0]), np.exp(x[1]), test, reference)

    def g(x):
        gu, gw = gradient_log(np.exp(x[0]), np.exp(x[1]), test, reference)
        return -np.array([gu, gw])

    scale = abs(f(x0))
    r = optimize.minimize(lambda x: f(x) / scale, x0, jac=lambda x: g(x) / scale, method="BFGS", options=dict(gtol=1e-13, maxiter=500))
    r.fun *= scale
    a, b = np.exp(r.x)
    return a / (a + b), 1.0 / (a + b + 1.0), -r.fun


def get_power_betabinom(size, phi, p, alt_p):
    """R/tools.R:128-166, theory = FALSE, limit = FALSE (VGAM::dbetabinom.ab restated through betaln / gammaln)."""
    size = int(size)
    x = np.arange(size + 1.0)
    lchoose = special.gammaln(size + 1) - special.gammaln(x + 1) - special.gammaln(size - x + 1)

    def ld(a, b):
        return lchoose + special.betaln(a + x, b + size - x) - special.betaln(a, b)

    a0, b0 = p * (1 - phi) / phi, (1 - p) * (1 - phi) / phi
    a1, b1 = alt_p * (1 - phi) / phi, (1 - alt_p) * (1 - phi) / phi
    l1, l0 = ld(a1, b1), ld(a0, b0)
    return float(np.sum(np.exp(l1) * (np.log10(np.e) * (l1 - l0))))
