"""ctypes wrapper over oracle/oracle.c (oracle/_ref/liboracle_port.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_port.so")

_lib = None
_D = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_I = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.edo_lnbeta.restype = C.c_double
        L.edo_lnbeta.argtypes = [C.c_double, C.c_double]
        L.edo_gammastar.restype = C.c_double
        L.edo_gammastar.argtypes = [C.c_double]
        L.edo_log1plusx.restype = C.c_double
        L.edo_log1plusx.argtypes = [C.c_double]
        L.edo_lngamma_sgn.restype = C.c_double
        L.edo_lngamma_sgn.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.edo_error_events.restype = C.c_long
        L.edo_error_events.argtypes = [C.c_int]
        L.edo_emission.restype = None
        L.edo_emission.argtypes = [_D, _D, _I, _I, C.c_int64, C.c_int32, _D, _D]
        L.edo_get_loglike_matrix.restype = None
        L.edo_get_loglike_matrix.argtypes = [_D, _D, _I, _I, C.c_double, C.c_int64, _D]
        L.edo_hmm.restype = C.c_int
        L.edo_hmm.argtypes = [C.c_int32, C.c_int32, _D, _D, _I, C.c_double, _I, _I, C.POINTER(C.c_int32)]
        L.edo_log_transition_table.restype = None
        L.edo_log_transition_table.argtypes = [C.c_int32, C.c_int32, _D, _I, C.c_double, _D]
        L.edo_forward_loglik.restype = C.c_double
        L.edo_forward_loglik.argtypes = [C.c_int32, C.c_int32, _D, _D, _I, C.c_double]
        L.edo_callcnvs_transitions.restype = None
        L.edo_callcnvs_transitions.argtypes = [C.c_int32, C.c_double, _D]
        _lib = L
    return _lib


def _d(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


def _i(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))


def lnbeta(x, y):
    f = lib().edo_lnbeta
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    return np.array([f(a, b) for a, b in zip(x.ravel(), y.ravel())]).reshape(x.shape)


def gammastar(x):
    f = lib().edo_gammastar
    x = np.asarray(x, float)
    return np.array([f(a) for a in x.ravel()]).reshape(x.shape)


def log1plusx(x):
    f = lib().edo_log1plusx
    x = np.asarray(x, float)
    return np.array([f(a) for a in x.ravel()]).reshape(x.shape)


def lngamma_sgn(x):
    f = lib().edo_lngamma_sgn
    x = np.asarray(x, float)
    s = C.c_double()
    vals, sgns = [], []
    for a in x.ravel():
        vals.append(f(a, C.byref(s)))
        sgns.append(s.value)
    return np.array(vals).reshape(x.shape), np.array(sgns).reshape(x.shape)


def error_events(reset=True):
    return lib().edo_error_events(1 if reset else 0)


def reference_odds(mixture=1.0):
    """CNV_estimate.cpp:65-66."""
    return np.array([1 - 0.5 * mixture, 1.0, 1 + 0.5 * mixture])


def state_odds(n_states, mixture=1.0, cn0_floor=0.05):
    """S-state odds table (extension for S>3, identical to the reference at S=3).

    S=3: CN {1,2,3}; S=5: CN {0,1,2,3,4}; S=7: CN {0..6}.  odds = 1 + (CN-2)/2*mixture, the CN0
    entry floored at `cn0_floor`.  Column order is by copy number (so S=3 stays del, normal, dup).
    """
    if n_states == 3:
        return reference_odds(mixture)
    cn = np.arange(0, n_states, dtype=np.float64)
    odds = 1 + (cn - 2) / 2 * mixture
    odds[2] = 1.0
    odds[0] = max(odds[0], cn0_floor)
    return odds


def emission(phi, expected, total, observed, odds):
    """n × S matrix, columns in `odds` order."""
    total, observed = _i(total), _i(observed)
    n = total.size
    phi = _d(np.broadcast_to(np.asarray(phi, float), (n,)))
    expected = _d(np.broadcast_to(np.asarray(expected, float), (n,)))
    odds = _d(odds)
    out = np.empty(n * odds.size)
    lib().edo_emission(phi, expected, total, observed, n, odds.size, odds, out)
    return out.reshape((n, odds.size), order="F")


def get_loglike_matrix(phi, expected, total, observed, mixture=1.0):
    total, observed = _i(total), _i(observed)
    n = total.size
    out = np.empty(n * 3)
    lib().edo_get_loglike_matrix(_d(np.broadcast_to(np.asarray(phi, float), (n,))),
                                 _d(np.broadcast_to(np.asarray(expected, float), (n,))),
                                 total, observed, float(mixture), n, out)
    return out.reshape((n, 3), order="F")


def c_hmm(transitions, loglikelihood, positions, expected_length):
    T = np.asarray(transitions, float)
    S = T.shape[0]
    ll = np.asarray(loglikelihood, float)
    nobs = ll.shape[0]
    path = np.zeros(nobs, np.int32)
    calls = np.zeros(4 * max(nobs, 1), np.int32)
    nc = C.c_int32(0)
    rc = lib().edo_hmm(S, nobs, _d(T.ravel(order="F")), _d(ll.ravel(order="F")),
                       _i(positions), float(expected_length), path, calls, C.byref(nc))
    if rc:
        return None
    return path, calls[:4 * nc.value].reshape(-1, 4).astype(np.int64)


def log_transition_table(transitions, positions, expected_length):
    T = np.asarray(transitions, float)
    S = T.shape[0]
    pos = _i(positions)
    out = np.empty(pos.size * S * S)
    lib().edo_log_transition_table(S, pos.size, _d(T.ravel(order="F")), pos, float(expected_length), out)
    return out.reshape(pos.size, S, S)  # [i][j][k]


def forward_loglik(transitions, loglikelihood, positions, expected_length):
    T = np.asarray(transitions, float)
    ll = np.asarray(loglikelihood, float)
    return lib().edo_forward_loglik(T.shape[0], ll.shape[0], _d(T.ravel(order="F")), _d(ll.ravel(order="F")),
                                    _i(positions), float(expected_length))


def callcnvs_transitions(n_states, tp):
    out = np.empty(n_states * n_states)
    lib().edo_callcnvs_transitions(n_states, float(tp), out)
    return out.reshape((n_states, n_states), order="F")


def tp_grid_loglik(ll_rows, offsets, start, end, tp_grid, expected_length=50000.0):
    """Extension (no reference counterpart): forward log-likelihood of one sample summed over chromosomes for every
    transition probability of the grid, with the CallCNVs framing.  ll_rows: n_bins x S in likelihood-column order.
    Returns (loglik[n_grid], index of the first maximiser)."""
    from . import framing
    S = ll_rows.shape[1]
    out = np.zeros(len(tp_grid))
    for gi, tp in enumerate(tp_grid):
        T = callcnvs_transitions(S, tp)
        tot = 0.0
        for c in range(len(offsets) - 1):
            b0, b1 = int(offsets[c]), int(offsets[c + 1])
            loc, pos = framing.frame_chromosome(ll_rows[b0:b1], np.asarray(start[b0:b1], float), np.asarray(end[b0:b1], float),
                                                expected_length)
            tot += forward_loglik(T, loc, pos, expected_length)
        out[gi] = tot
    return out, int(np.argmax(out))
