"""The reference's own compiled C/C++ hot path (oracle/_ref/libexomedepth_ref.so).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Built by `make -C oracle ref` from the sources
where they lie under /root/reference/src; never read at run time from /root/reference.
"""
import ctypes as C
import os

import numpy as np

from . import sexp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libexomedepth_ref.so")


def available():
    return os.path.exists(LIB_PATH)


_api = None


def api():
    global _api
    if _api is None:
        lib = C.CDLL(LIB_PATH)
        lib.gsl_sf_lnbeta.restype = C.c_double
        lib.gsl_sf_lnbeta.argtypes = [C.c_double, C.c_double]
        lib.gsl_sf_lngamma.restype = C.c_double
        lib.gsl_sf_lngamma.argtypes = [C.c_double]
        lib.gsl_sf_gammastar.restype = C.c_double
        lib.gsl_sf_gammastar.argtypes = [C.c_double]
        lib.gsl_sf_log_1plusx.restype = C.c_double
        lib.gsl_sf_log_1plusx.argtypes = [C.c_double]
        _api = sexp.CallApi(lib)
    return _api


def get_loglike_matrix(phi, expected, total, observed, mixture=1.0):
    return api().get_loglike_matrix(phi, expected, total, observed, mixture)


def c_hmm(transitions, loglikelihood, positions, expected_length, nstates=None):
    return api().c_hmm(transitions, loglikelihood, positions, expected_length, nstates)


def lnbeta(x, y):
    """src/beta.c:161-164."""
    f = api().lib.gsl_sf_lnbeta
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    return np.array([f(a, b) for a, b in zip(x.ravel(), y.ravel())]).reshape(x.shape)


def lngamma(x):
    f = api().lib.gsl_sf_lngamma
    x = np.asarray(x, float)
    return np.array([f(a) for a in x.ravel()]).reshape(x.shape)


def gammastar(x):
    f = api().lib.gsl_sf_gammastar
    x = np.asarray(x, float)
    return np.array([f(a) for a in x.ravel()]).reshape(x.shape)


def log1plusx(x):
    f = api().lib.gsl_sf_log_1plusx
    x = np.asarray(x, float)
    return np.array([f(a) for a in x.ravel()]).reshape(x.shape)
