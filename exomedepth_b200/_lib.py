"""ctypes binding of libexomedepth_b200.so (the C ABI of include/exomedepth_b200.h).

The library is CUDA-only.  If it is missing, or no sm_100 device is usable, every entry point raises:
there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EDB200_LIB: another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("EDB200_LIB") or os.path.join(_HERE, "libexomedepth_b200.so")

OK, WARN_NAN, ERR_NSTATES, ERR_CUDA, ERR_ARG, WARN_CALLCAP = 0, 1, 2, 4, 8, 16
MAX_STATES = 7
EMISSION_AUTO, EMISSION_DIRECT, EMISSION_TABLE, EMISSION_PANEL = 0, 1, 2, 3
OPTIONS = dict(sweep=1, parts=2, vsplit=3, crit_warps=4, sweep_warps=5, packplan=6, segments=7, seg_warm=8, seg_min=9, seg_repair=10, reserve=11, chunks=12)      # EDB200_OPT_*
SWEEP_AUTO, SWEEP_LANE_PER_STATE, SWEEP_THREAD_PER_CHAIN = 0, 1, 2

EXPORTS = (
    "edb200_init", "edb200_shutdown", "edb200_last_error", "edb200_device_info", "edb200_launch_count",
    "edb200_host_alloc", "edb200_host_free", "edb200_pack_counts16", "edb200_pack_counts12", "edb200_get_loglike_matrix", "edb200_emission", "edb200_lnbeta", "edb200_hmm",
    "edb200_cohort_create", "edb200_cohort_destroy", "edb200_cohort_set_option", "edb200_cohort_segment_stats", "edb200_cohort_table", "edb200_cohort_table_copy",
    "edb200_cohort_run_device", "edb200_cohort_capture_device", "edb200_graph_launch", "edb200_graph_destroy",
    "edb200_cohort_run_host", "edb200_status", "edb200_profile", "edb200_profile_read",
    "edb200_cohort_forward_device", "edb200_cohort_forward_last",
    "edb200_refset_correlations", "edb200_refset_kpad", "edb200_refset_standardize_device", "edb200_refset_gram_device",
    "edb200_betabin_fit", "edb200_betabin_fit_device", "edb200_power_betabinom", "edb200_gsl_error_log",
    "edb200_refset_block_alloc", "edb200_refset_peers_open", "edb200_refset_peers_close", "edb200_refset_gram_peers_device",
)


class EDB200Error(RuntimeError):
    pass


class CohortSpec(C.Structure):
    _fields_ = [("n_bins", C.c_int64), ("n_chains", C.c_int32), ("chain_offsets", C.c_void_p),
                ("start", C.c_void_p), ("end", C.c_void_p), ("n_states", C.c_int32), ("odds", C.c_void_p),
                ("mixture", C.c_double), ("transitions", C.c_void_p), ("transition_probability", C.c_double),
                ("expected_cnv_length", C.c_double), ("skip_table_build", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("observed", C.c_void_p), ("obs_stride", C.c_int64),
                ("reference", C.c_void_p), ("ref_stride", C.c_int64), ("phi", C.c_void_p),
                ("expected", C.c_void_p), ("ll", C.c_void_p), ("ll_stride", C.c_int64), ("path", C.c_void_p),
                ("path_stride", C.c_int64), ("calls", C.c_void_p), ("ncalls", C.c_void_p), ("call_cap", C.c_int32),
                ("call_stats", C.c_void_p), ("cor", C.c_void_p),
                ("per_bin_stride", C.c_int64), ("observed16", C.c_void_p), ("obs16_stride", C.c_int64), ("n_overflow", C.c_int64),
                ("overflow_index", C.c_void_p), ("overflow_value", C.c_void_p),
                ("observed12", C.c_void_p), ("obs12_stride", C.c_int64)]


_lib = None


def load():
    """dlopen the library and declare prototypes.  Does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EDB200Error(f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C exomedepth_b200/csrc`); exomedepth_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.edb200_init.restype = C.c_int
    L.edb200_init.argtypes = [C.c_int]
    L.edb200_shutdown.restype = None
    L.edb200_last_error.restype = C.c_char_p
    L.edb200_device_info.restype = C.c_int
    L.edb200_device_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.edb200_launch_count.restype = i64
    L.edb200_launch_count.argtypes = [C.c_int]
    L.edb200_host_alloc.restype = vp
    L.edb200_host_alloc.argtypes = [C.c_size_t]
    L.edb200_host_free.restype = None
    L.edb200_host_free.argtypes = [vp]
    for fn in (L.edb200_pack_counts16, L.edb200_pack_counts12):
        fn.restype = i64
        fn.argtypes = [vp, i64, i32, i64, vp, i64, vp, vp, i64]
    L.edb200_get_loglike_matrix.restype = C.c_int
    L.edb200_get_loglike_matrix.argtypes = [vp, vp, vp, vp, dbl, i64, vp]
    L.edb200_lnbeta.restype = C.c_int
    L.edb200_lnbeta.argtypes = [vp, vp, i64, vp]
    L.edb200_emission.restype = C.c_int
    L.edb200_emission.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
    L.edb200_hmm.restype = C.c_int
    L.edb200_hmm.argtypes = [i32, i32, vp, vp, vp, dbl, vp, vp, i32, vp]
    L.edb200_cohort_create.restype = C.c_int
    L.edb200_cohort_create.argtypes = [C.POINTER(CohortSpec), C.POINTER(vp)]
    L.edb200_cohort_destroy.restype = None
    L.edb200_cohort_destroy.argtypes = [vp]
    L.edb200_cohort_set_option.restype = C.c_int
    L.edb200_cohort_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.edb200_cohort_segment_stats.restype = C.c_int
    L.edb200_cohort_segment_stats.argtypes = [vp, C.POINTER(C.c_int32)]
    L.edb200_cohort_table.restype = C.c_int
    L.edb200_cohort_table.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.edb200_cohort_table_copy.restype = C.c_int
    L.edb200_cohort_table_copy.argtypes = [vp, vp, C.c_int, vp]
    L.edb200_cohort_run_device.restype = C.c_int
    L.edb200_cohort_run_device.argtypes = [vp, C.POINTER(Batch), C.c_int, C.c_int, vp]
    L.edb200_cohort_capture_device.restype = C.c_int
    L.edb200_cohort_capture_device.argtypes = [vp, C.POINTER(Batch), C.c_int, C.c_int, C.POINTER(vp)]
    L.edb200_graph_launch.restype = C.c_int
    L.edb200_graph_launch.argtypes = [vp, vp]
    L.edb200_graph_destroy.restype = None
    L.edb200_graph_destroy.argtypes = [vp]
    L.edb200_cohort_run_host.restype = C.c_int
    L.edb200_cohort_run_host.argtypes = [vp, C.POINTER(Batch), C.c_int]
    L.edb200_cohort_forward_device.restype = C.c_int
    L.edb200_cohort_forward_device.argtypes = [vp, C.POINTER(Batch), vp, i32, vp, vp, vp]
    L.edb200_cohort_forward_last.restype = C.c_int
    L.edb200_cohort_forward_last.argtypes = [vp, vp, i32, vp, vp]
    L.edb200_refset_correlations.restype = C.c_int
    L.edb200_refset_correlations.argtypes = [vp, i64, i32, vp, vp, i64, i32, i32, vp]
    L.edb200_refset_kpad.restype = i64
    L.edb200_refset_kpad.argtypes = [i64]
    L.edb200_refset_standardize_device.restype = C.c_int
    L.edb200_refset_standardize_device.argtypes = [vp, i64, i32, vp, vp, i64, vp, vp]
    L.edb200_refset_gram_device.restype = C.c_int
    L.edb200_refset_gram_device.argtypes = [vp, i32, vp, i32, i64, vp, vp]
    L.edb200_refset_block_alloc.restype = C.c_int
    L.edb200_refset_block_alloc.argtypes = [i32, i64, C.POINTER(vp), vp]
    L.edb200_refset_peers_open.restype = C.c_int
    L.edb200_refset_peers_open.argtypes = [vp, i32, i32]
    L.edb200_refset_peers_close.restype = C.c_int
    L.edb200_refset_peers_close.argtypes = []
    L.edb200_refset_gram_peers_device.restype = C.c_int
    L.edb200_refset_gram_peers_device.argtypes = [i32, i32, i32, i64, vp, vp]
    L.edb200_betabin_fit.restype = C.c_int
    L.edb200_betabin_fit.argtypes = [vp, i64, vp, i64, i32, i64, vp, vp, vp, vp]
    L.edb200_gsl_error_log.restype = C.c_int64
    L.edb200_gsl_error_log.argtypes = [C.c_char_p, C.c_size_t, C.c_int64, C.POINTER(C.c_int64)]
    L.edb200_power_betabinom.restype = C.c_int
    L.edb200_power_betabinom.argtypes = [vp, vp, vp, vp, i32, vp]
    L.edb200_betabin_fit_device.restype = C.c_int
    L.edb200_betabin_fit_device.argtypes = [vp, i64, vp, i64, i32, i64, vp, vp, vp, vp, vp]
    L.edb200_status.restype = C.c_int
    L.edb200_status.argtypes = [C.c_int]
    L.edb200_profile.restype = C.c_int
    L.edb200_profile.argtypes = [C.c_int]
    L.edb200_profile_read.restype = C.c_int
    L.edb200_profile_read.argtypes = [C.c_char_p, C.c_int]
    _lib = L
    return L


def last_error():
    return load().edb200_last_error().decode()


def check(rc, what):
    """Raise on failure codes; return the warning bits (WARN_NAN, WARN_CALLCAP)."""
    if rc & (ERR_NSTATES | ERR_CUDA | ERR_ARG):
        raise EDB200Error(f"{what}: {last_error()} (status {rc})")
    return rc


def init(device=-1):
    check(load().edb200_init(int(device)), "edb200_init")


def device_info():
    buf = C.create_string_buffer(256)
    sms, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    check(load().edb200_device_info(buf, 256, C.byref(sms), C.byref(maj), C.byref(mnr)), "edb200_device_info")
    return dict(name=buf.value.decode(), n_sms=sms.value, cc=(maj.value, mnr.value))


def launch_count(reset=False):
    return int(load().edb200_launch_count(1 if reset else 0))


def profile(enable):
    """Bracket every kernel launch with CUDA events on its stream (bench.py's per-kernel roofline)."""
    check(load().edb200_profile(int(enable)), "edb200_profile")      # 2: also print the launch timeline on read


def profile_read():
    """{kernel: (launches, total_ms, longest_ms)} since profiling was enabled or last read."""
    buf = C.create_string_buffer(4096)
    check(load().edb200_profile_read(buf, 4096), "edb200_profile_read")
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, n, ms, mx = item.split(":")
            out[name] = (int(n), float(ms), float(mx))
    return out


class PinnedPool:
    """numpy views over page-locked host buffers from edb200_host_alloc (freed when the pool is closed)."""

    def __init__(self):
        self._ptrs = []

    def empty(self, shape, dtype):
        import numpy as np
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = load().edb200_host_alloc(max(n, 1))
        if not p:
            raise EDB200Error(f"edb200_host_alloc({n}): {last_error()}")
        self._ptrs.append(p)
        buf = (C.c_char * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        for p in self._ptrs:
            load().edb200_host_free(p)
        self._ptrs = []
