"""Beta-binomial fit of `new('ExomeDepth')` on the GPU (stand-in for aod::betabin with the default formula
`cbind(test, reference) ~ 1`, R/class_definition.R:118-119, 168) and the expected Bayes factor of
`get.power.betabinom` (R/tools.R:128-166, the deterministic beta-binomial branch)."""
import math

import numpy as np

from . import _lib

INFO = {-1: "degenerate sample (no reads, or every read in the test)", -2: "a per-bin count exceeds the fit kernel's caps",
        -3: "iteration cap reached", -4: "no over-dispersion: phi -> 0 (binomial limit)"}


def fit(observed, reference):
    """observed: int32[n_samples, n_bins] (or one vector); reference: int32[n_bins] shared by all samples, or one row
    per sample.  Returns dict(expected, phi, loglik, info) with one entry per sample: the maximum-likelihood expected
    proportion (x@expected, constant over the bins), over-dispersion (x@phi), maximised log-likelihood (without the
    binomial coefficients) and Newton iterations (negative: see INFO)."""
    obs = np.ascontiguousarray(np.atleast_2d(np.asarray(observed, np.int32)))
    ref = np.ascontiguousarray(np.asarray(reference, np.int32))
    ns, nb = obs.shape
    if ref.shape not in ((nb,), (ns, nb)):
        raise ValueError("Length of test and numeric must match")                # R/class_definition.R:92
    mu, phi, ll = np.empty(ns), np.empty(ns), np.empty(ns)
    info = np.empty(ns, np.int32)
    rc = _lib.load().edb200_betabin_fit(obs.ctypes.data, nb, ref.ctypes.data, 0 if ref.ndim == 1 else nb, ns, nb,
                                        mu.ctypes.data, phi.ctypes.data, ll.ctypes.data, info.ctypes.data)
    _lib.check(rc, "edb200_betabin_fit")
    return dict(expected=mu, phi=phi, loglik=ll, info=info)


def get_power_betabinom_batch(size, my_phi, my_p, my_alt_p):
    """get.power.betabinom for arrays of problems on the GPU (edb200_power_betabinom, one CTA per problem); the scalar
    function below is the host restatement the tests compare it with."""
    size = np.ascontiguousarray(np.asarray(size, np.int32))
    phi, p, ap = (np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), size.shape)) for v in (my_phi, my_p, my_alt_p))
    out = np.empty(size.shape, np.float64)
    _lib.check(_lib.load().edb200_power_betabinom(size.ctypes.data, phi.ctypes.data, p.ctypes.data, ap.ctypes.data, size.size, out.ctypes.data),
               "edb200_power_betabinom")
    return out


def _ldbetabinom(x, size, a, b):
    """log dbetabinom.ab(x, size, a, b) — VGAM's density, from lgamma."""
    lg = math.lgamma
    return (lg(size + 1) - lg(x + 1) - lg(size - x + 1) + lg(a + x) + lg(b + size - x) - lg(a + b + size)
            - (lg(a) + lg(b) - lg(a + b)))


def get_power_betabinom(size, my_phi, my_p, my_alt_p):
    """R/tools.R:128-166 with theory = FALSE, limit = FALSE: the expected log10 Bayes factor of the alternative
    proportion against the null over one beta-binomial draw of `size` reads."""
    size = int(size)
    a0, b0 = my_p * (1 - my_phi) / my_phi, (1 - my_p) * (1 - my_phi) / my_phi
    a1, b1 = my_alt_p * (1 - my_phi) / my_phi, (1 - my_alt_p) * (1 - my_phi) / my_phi
    log10e = math.log10(math.e)
    total = 0.0
    terms = []
    for x in range(size + 1):
        l1, l0 = _ldbetabinom(x, size, a1, b1), _ldbetabinom(x, size, a0, b0)
        terms.append(math.exp(l1) * (log10e * (l1 - l0)))
    total = math.fsum(terms)
    return total
