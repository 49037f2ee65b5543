"""Host-side mirror of the reference's operator interface for the hot path (names, argument meaning and
error behaviour follow the reference; the arithmetic runs on the GPU through the C ABI)."""
import math
import sys

import numpy as np

from . import _lib


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


def _i32(x):
    a = np.asarray(x)
    if a.dtype.kind == "f":
        a = np.trunc(a)            # as.integer()
    return np.ascontiguousarray(a.astype(np.int32).reshape(-1))


def _warn_nan(rc, what):
    if rc & _lib.WARN_NAN:
        # the reference prints through Rprintf and carries on (src/error.c:45-48)
        print(f"ERROR {what}: domain error\nDefault GSL error handler invoked.", file=sys.stderr)


def get_loglike_matrix(phi, expected, total, observed, mixture=1.0):
    """`.Call("get_loglike_matrix", phi, expected, total, observed, mixture)` — src/CNV_estimate.cpp:52-85.

    Returns the n×3 matrix (columns deletion, normal, duplication)."""
    total, observed = _i32(total), _i32(observed)
    n = total.size
    phi = _f64(np.broadcast_to(np.asarray(phi, float), (n,)))
    expected = _f64(np.broadcast_to(np.asarray(expected, float), (n,)))
    if mixture != 1:
        print("As a warning (this could be normal), the mixture coefficient is %f" % mixture, file=sys.stderr)  # :61
    out = np.empty(n * 3)
    rc = _lib.load().edb200_get_loglike_matrix(phi.ctypes.data, expected.ctypes.data, total.ctypes.data,
                                               observed.ctypes.data, float(mixture), n, out.ctypes.data)
    _warn_nan(_lib.check(rc, "get_loglike_matrix"), "get_loglike_matrix")
    return out.reshape((n, 3), order="F")


def lnbeta(x, y):
    """gsl_sf_lnbeta (src/beta.c:161-164) element-wise through the device's restatement of the vendored GSL chain
    (edb200_lnbeta).  NaN where the reference raises a domain error."""
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    shape = x.shape
    x, y = _f64(x.ravel()), _f64(y.ravel())
    out = np.empty(x.size)
    _lib.check(_lib.load().edb200_lnbeta(x.ctypes.data, y.ctypes.data, x.size, out.ctypes.data), "edb200_lnbeta")
    return out.reshape(shape)


def emission(phi, expected, total, observed, odds):
    """S-state generalisation (extension): columns in the order of `odds`."""
    total, observed = _i32(total), _i32(observed)
    n = total.size
    phi = _f64(np.broadcast_to(np.asarray(phi, float), (n,)))
    expected = _f64(np.broadcast_to(np.asarray(expected, float), (n,)))
    odds = _f64(odds)
    out = np.empty(n * odds.size)
    rc = _lib.load().edb200_emission(phi.ctypes.data, expected.ctypes.data, total.ctypes.data, observed.ctypes.data,
                                     n, odds.size, odds.ctypes.data, out.ctypes.data)
    _warn_nan(_lib.check(rc, "emission"), "emission")
    return out.reshape((n, odds.size), order="F")


def C_hmm(nstates, nobs, transitions, loglikelihood, positions, expected_length, strict_reference=True):
    """`.Call("C_hmm", nstates, nobs, transitions, loglikelihood, positions, expectedLength)` — src/hmm.cpp:18-167.

    Returns (path int32[nobs], calls int64[ncalls, 4]) — or None after printing the reference's message when
    nstates != 3 and strict_reference is set (hmm.cpp:37-40)."""
    if strict_reference and nstates != 3:
        print("ERROR: The code must assume 3 states", file=sys.stderr)
        return None
    T = np.asfortranarray(np.asarray(transitions, float))
    ll = np.asfortranarray(np.asarray(loglikelihood, float))
    pos = _i32(positions)
    path = np.empty(nobs, np.int32)
    cap = max(nobs, 1)          # every state change can close a call (hmm.cpp:110-121)
    calls = np.empty(4 * cap, np.int32)
    import ctypes as C
    nc = C.c_int32(0)
    rc = _lib.load().edb200_hmm(int(nstates), int(nobs), T.ctypes.data, ll.ctypes.data, pos.ctypes.data,
                                float(expected_length), path.ctypes.data, calls.ctypes.data, cap, C.addressof(nc))
    _lib.check(rc, "C_hmm")
    return path, calls[:4 * nc.value].reshape(-1, 4).astype(np.int64)


def viterbi_hmm(transitions, loglikelihood, positions, expected_CNV_length, strict_reference=True):
    """R/tools.R:88-103 — argument checks, then C_hmm; returns dict(Viterbi_path, calls)."""
    T = np.asarray(transitions, float)
    ll = np.asarray(loglikelihood, float)
    if T.shape[0] != T.shape[1]:
        raise ValueError("Transition matrix is not square")
    if len(positions) != ll.shape[0]:
        raise ValueError("The number of positions are not matching the number of rows of the likelihood matrix "
                         f"{len(positions)} and {ll.shape[0]}")
    res = C_hmm(T.shape[0], ll.shape[0], T, ll, positions, float(expected_CNV_length), strict_reference)
    if res is None:
        return None
    return dict(Viterbi_path=res[0], calls=res[1])


def _signif(x, digits=3):
    x = float(x)
    if x == 0 or not math.isfinite(x):
        return x
    return round(x, digits - 1 - math.floor(math.log10(abs(x))))


class ExomeDepth:
    """The S4 class of R/class_definition.R:36-46, restricted to the slots the hot path touches.

    `phi` and `expected` may be given (e.g. from an R-side `aod::betabin` fit); when they are None the beta-binomial
    model of the default formula `cbind(test, reference) ~ 1` is fitted on the GPU (betabin.py; the stand-in for
    aod::betabin, R/class_definition.R:118-119, 168 — parity unpinned, the result is the likelihood maximiser)."""

    def __init__(self, test, reference, phi=None, expected=None, prop_tumor=1.0, verbose=False, positions=None):
        test = np.asarray(test, float)
        reference = np.asarray(reference, float)
        if test.size != reference.size:
            raise ValueError("Length of test and numeric must match")
        self.test, self.reference = test, reference
        # `positions` (the GRanges slot, R/class_definition.R:44): dict(chromosome, start, end), one entry per bin
        if positions is not None and len(positions["start"]) != test.size:
            raise ValueError("The provided genomic positions (GRanges object) and test count vector are not matching in length: "
                             f"{len(positions['start'])}, {test.size}")                       # R/class_definition.R:163-165
        self.positions = positions
        self.likelihood = None
        self.annotations = None
        self.CNV_calls = []
        self.cor_test_reference = float("nan")
        if np.sum(test > 5) < 5:                      # R/class_definition.R:95-98
            self.phi = np.zeros(0)
            self.expected = np.zeros(0)
            return
        if phi is None or expected is None:
            from . import betabin
            if verbose:
                print(f"Now fitting the beta-binomial model on a data frame with {test.size} rows : this step can take a few minutes.",
                      file=sys.stderr)
            fit = betabin.fit(_i32(test), _i32(reference))
            if fit["info"][0] in (-1, -2):
                raise _lib.EDB200Error(f"beta-binomial fit failed: {betabin.INFO[int(fit['info'][0])]}")
            phi = fit["phi"][0] if phi is None else phi
            expected = fit["expected"][0] if expected is None else expected
        self.phi = np.broadcast_to(np.asarray(phi, float), test.shape).copy()
        self.expected = np.broadcast_to(np.asarray(expected, float), test.shape).copy()
        if verbose:
            print("Now computing the likelihood for the different copy number states", file=sys.stderr)
        self.likelihood = get_loglike_matrix(self.phi, self.expected, _i32(reference + test), _i32(test), prop_tumor)


def TestCNV(x, chromosome, start, end, type):
    """R/class_definition.R:243-256 — the log likelihood ratio in favour of a CNV of the given type over the bins that
    lie inside [start, end] on `chromosome` (the likelihood matrix is the one the GPU computed at construction)."""
    if type not in ("deletion", "duplication"):
        raise ValueError("type must be either duplication or deletion")
    if not isinstance(chromosome, str):
        raise ValueError("The input chromosome must be a character or a factor")
    if getattr(x, "positions", None) is None:
        raise ValueError("This function cannot be used if the position of the exons/DNA segments was not included in the ExomeDepth object")
    pos = x.positions
    inside = (np.asarray([str(c) for c in pos["chromosome"]]) == chromosome) & (np.asarray(pos["start"]) >= start) & (np.asarray(pos["end"]) <= end)
    col = 0 if type == "deletion" else 2
    return float(math.fsum(x.likelihood[inside, col] - x.likelihood[inside, 1]))


def somatic_CNV_call(normal, tumor, prop_tumor=1.0, chromosome=None, start=None, end=None, names=None):
    """R/class_definition.R:442-461 — tumour against its matched normal: the tumour is the test sample, the normal the
    reference, `prop.tumor` the mixture coefficient of the emission model; beta-binomial fit and CallCNVs as usual."""
    print("Warning: this function is largely untested and experimental", file=sys.stderr)       # :444
    x = ExomeDepth(tumor, normal, prop_tumor=prop_tumor)
    return CallCNVs(x, chromosome, start, end, names, transition_probability=1e-4)


def CallCNVs(x, chromosome, start, end, name, transition_probability=1e-4, expected_CNV_length=50000):
    """R/class_definition.R:311-419 on top of the GPU Viterbi. Fills and returns x (x.CNV_calls: list of dict)."""
    if x.phi.size == 0:
        x.CNV_calls = []
        return x
    n = len(chromosome)
    if len(start) != n or len(end) != n or len(name) != n:
        raise ValueError("Chromosome, name, start and end vector must have the same lengths.")
    if x.likelihood.shape[0] != n:
        raise ValueError("The annotation vectors must have the same length as the data in the ExomeDepth x")
    used = list(dict.fromkeys(str(c) for c in chromosome))
    auto = [str(i) for i in range(1, 23)]
    levels = [c for c in auto + [c for c in used if c not in auto] if c in used]
    code = np.array([levels.index(str(c)) for c in chromosome])
    start = np.asarray(start, float)
    end = np.asarray(end, float)
    name = np.asarray(name, dtype=object)
    order = np.lexsort((0.5 * (start + end), code))
    if np.any(order != np.arange(n)):
        x.test, x.reference, x.likelihood = x.test[order], x.reference[order], x.likelihood[order]
        code, start, end, name = code[order], start[order], end[order], name[order]
    x.annotations = dict(name=name, chromosome=code, levels=levels, start=start, end=end)
    x.cor_test_reference = float(np.corrcoef(x.test, x.reference)[0, 1])
    total = x.test + x.reference
    tp = transition_probability
    T = np.array([[1. - tp, tp / 2., tp / 2.], [0.5, 0.5, 0.], [0.5, 0., 0.5]])
    final, shift = [], 0
    for c in dict.fromkeys(code.tolist()):
        good = np.nonzero(code == c)[0]
        loc_ll = x.likelihood[good]
        loglik = np.vstack([[-np.inf, 0, -np.inf], loc_ll[:, [1, 0, 2]], [-100, 0, -100]])
        pos = np.trunc(np.concatenate([[start[good][0] - 2 * expected_CNV_length], start[good],
                                       [end[good][-1] + 2 * expected_CNV_length]])).astype(np.int32)
        res = viterbi_hmm(T, loglik, pos, expected_CNV_length)
        for sp, ep, typ, nex in res["calls"]:
            sp, ep = int(sp) - 1, int(ep) - 1
            sl = slice(sp - 1, ep)
            col = {1: 0, 2: 2}[int(typ)]
            bf = float(np.sum(loc_ll[sl, col] - loc_ll[sl, 1]))
            rexp = int(np.sum(total[good][sl] * x.expected[good][sl]))
            robs = float(np.sum(x.test[good][sl]))
            chrom = levels[c]
            ident = f"chr{chrom}:{int(start[good][sp - 1])}-{int(end[good][ep - 1])}".replace("chrchr", "chr")
            final.append(dict(start_p=sp + shift, end_p=ep + shift, type=("deletion", "duplication")[int(typ) - 1],
                              nexons=int(nex), start=float(start[good][sp - 1]), end=float(end[good][ep - 1]),
                              chromosome=chrom, id=ident, BF=_signif(math.log10(math.e) * bf, 3),
                              reads_expected=rexp, reads_observed=robs,
                              reads_ratio=_signif(robs / rexp, 3) if rexp else float("inf")))
        shift += good.size
    x.CNV_calls = final
    return x
