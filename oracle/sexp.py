"""Fake-SEXP plumbing for libraries compiled against oracle/stub/Rinternals.h.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Mirrors ``struct edb200_stub_sexp``.
The same helpers drive the compiled reference (``oracle/_ref/libexomedepth_ref.so``) and this
repo's own ``.Call`` glue built against the stand-in R API (``oracle/_ref/librglue_stub.so``),
so both are exercised the way R's ``.Call`` would: R/class_definition.R:184-189, R/tools.R:97.
"""
import ctypes as C

import numpy as np

INTSXP, REALSXP, VECSXP = 13, 14, 19


class SexpRec(C.Structure):
    _fields_ = [("type", C.c_int), ("n", C.c_int), ("nrow", C.c_int), ("ncol", C.c_int),
                ("data", C.c_void_p)]


SEXP = C.POINTER(SexpRec)


class Holder:
    """Keeps the numpy buffer alive for as long as the fake SEXP is in use."""

    def __init__(self, arr, sxp_type, nrow=0, ncol=0):
        self.arr = arr
        self.rec = SexpRec(sxp_type, int(arr.size), int(nrow), int(ncol), arr.ctypes.data)

    @property
    def ptr(self):
        return C.pointer(self.rec)


def real(x, matrix=False):
    a = np.asarray(x, dtype=np.float64)
    if matrix:
        a = np.asfortranarray(a)
        return Holder(a, REALSXP, a.shape[0], a.shape[1])
    return Holder(np.ascontiguousarray(a).reshape(-1), REALSXP)


def integer(x):
    return Holder(np.ascontiguousarray(np.asarray(x, dtype=np.int32)).reshape(-1), INTSXP)


def read_real(p):
    """Copy a REALSXP (vector or column-major matrix) out of a library-allocated SEXP."""
    rec = p.contents
    assert rec.type == REALSXP, rec.type
    if rec.n == 0:
        flat = np.zeros(0)
    else:
        flat = np.ctypeslib.as_array(C.cast(rec.data, C.POINTER(C.c_double)), shape=(rec.n,)).copy()
    if rec.nrow or rec.ncol:
        return flat.reshape((rec.nrow, rec.ncol), order="F")
    return flat


def list_elt(p, i):
    rec = p.contents
    assert rec.type == VECSXP
    arr = C.cast(rec.data, C.POINTER(SEXP))
    return arr[i]


def bind_call_api(lib):
    """Declare the two .Call routines registered in src/ExomeDepth_init.c:14-24."""
    lib.get_loglike_matrix.restype = SEXP
    lib.get_loglike_matrix.argtypes = [SEXP] * 5
    lib.C_hmm.restype = SEXP
    lib.C_hmm.argtypes = [SEXP] * 6
    lib.edb200_stub_free.restype = None
    lib.edb200_stub_free.argtypes = [SEXP]
    lib.edb200_stub_rprintf_count.restype = C.c_int
    lib.edb200_stub_rprintf_count.argtypes = [C.c_int]
    lib.edb200_stub_rprintf_quiet.restype = None
    lib.edb200_stub_rprintf_quiet.argtypes = [C.c_int]
    return lib


class CallApi:
    """`.Call("get_loglike_matrix", …)` / `.Call("C_hmm", …)` against any library exporting them."""

    def __init__(self, lib):
        self.lib = bind_call_api(lib)

    def get_loglike_matrix(self, phi, expected, total, observed, mixture=1.0):
        """src/CNV_estimate.cpp:52-85 — returns the n×3 matrix (columns del, normal, dup)."""
        h = [real(phi), real(expected), integer(total), integer(observed), real([mixture])]
        out = self.lib.get_loglike_matrix(*[x.ptr for x in h])
        res = read_real(out)
        self.lib.edb200_stub_free(out)
        return res

    def c_hmm(self, transitions, loglikelihood, positions, expected_length, nstates=None):
        """src/hmm.cpp:18-167 — returns (path int32[nobs], calls int64[ncalls,4]) or None.

        `loglikelihood` is nobs×S in HMM state order (0=normal, 1=deletion, 2=duplication).
        """
        ll = np.asarray(loglikelihood, dtype=np.float64)
        T = np.asarray(transitions, dtype=np.float64)
        ns = T.shape[0] if nstates is None else nstates
        h = [integer([ns]), integer([ll.shape[0]]), real(T, matrix=True), real(ll, matrix=True),
             integer(positions), real([expected_length])]
        out = self.lib.C_hmm(*[x.ptr for x in h])
        if not out:
            return None
        path = read_real(list_elt(out, 0)).astype(np.int32)
        calls = read_real(list_elt(out, 1))
        calls = calls.reshape(-1, 4).astype(np.int64) if calls.size else np.zeros((0, 4), np.int64)
        self.lib.edb200_stub_free(out)
        return path, calls

    def rprintf_count(self, reset=True):
        return self.lib.edb200_stub_rprintf_count(1 if reset else 0)

    def quiet(self, on=True):
        self.lib.edb200_stub_rprintf_quiet(1 if on else 0)
