"""Cohort: many samples over one shared, ordered bin set — the batched form of
`new('ExomeDepth')`'s likelihood step + `CallCNVs`' per-chromosome Viterbi (R/class_definition.R:184-189,
342-374).  torch is used only for device memory, streams and torch.distributed plumbing; all arithmetic
is in the CUDA kernels behind the C ABI."""
import ctypes as C

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _pack_native(bits, observed, out, out_stride):
    """edb200_pack_counts16 / edb200_pack_counts12 (host threads, include/exomedepth_b200.h) over an int32 matrix."""
    lib = _lib.load()
    fn = lib.edb200_pack_counts12 if bits == 12 else lib.edb200_pack_counts16
    ns, nb = observed.shape
    cap = 1 << 16
    while True:
        idx, val = np.empty(cap, np.int64), np.empty(cap, np.int32)
        n = fn(_ptr(observed), observed.strides[0] // 4, ns, nb, _ptr(out), out_stride, _ptr(idx), _ptr(val), cap)
        if n < 0:
            raise ValueError(lib.edb200_last_error().decode() or "pack_counts failed")
        if n <= cap:
            return idx[:n].copy(), val[:n].copy()
        cap = int(n)


def _as_count_matrix(observed):
    observed = np.asarray(observed)
    if observed.ndim != 2:
        raise ValueError("count matrix must be [n_samples, n_bins]")
    if observed.dtype != np.int32 or observed.strides[1] != 4 or observed.strides[0] % 4 or observed.strides[0] < 4 * observed.shape[1]:
        if observed.size and (observed.min() < 0 or observed.max() > np.iinfo(np.int32).max):
            raise ValueError("negative read count" if observed.min() < 0 else "read count beyond int32")
        observed = np.ascontiguousarray(observed, np.int32)
    return observed


def pack_counts(observed, out=None):
    """The 16-bit ingestion layout of a count matrix (include/exomedepth_b200.h, edb200_batch.observed16): returns
    (uint16[n_samples, n_bins] with 65535 = "see the overflow list", int64 flat indices sample * n_bins + bin, int32 values).
    `out`: an existing uint16 array (e.g. pinned) to fill.  This is the layout a loader writes ONCE per cohort
    (R/countBamInGranges.R:356-369 produces the counts); it halves the bytes every later call moves over PCIe.
    Encoded by the library's host-side encoder (edb200_pack_counts16); pack_counts_numpy is the same thing in numpy."""
    observed = _as_count_matrix(observed)
    u16 = out if out is not None else np.empty(observed.shape, np.uint16)
    assert u16.shape == observed.shape and u16.dtype == np.uint16 and u16.flags.c_contiguous
    idx, val = _pack_native(16, observed, u16, observed.shape[1])
    return u16, idx, val


def pack_counts12(observed, out=None):
    """The 12-bit ingestion layout (edb200_batch.observed12): returns (uint8[n_samples, stride] — every row a little-endian
    bit stream of 12 bits per bin, 4095 = "see the overflow list", stride = 3 * ceil(n_bins / 2) rounded up to 4 —, int64 flat
    indices sample * n_bins + bin, int32 values).  `out`: an existing uint8 array of that shape (e.g. pinned) to fill.  A
    quarter fewer bytes per call over PCIe than pack_counts.  Encoded by edb200_pack_counts12; pack_counts12_numpy is the same
    thing in numpy."""
    observed = _as_count_matrix(observed)
    ns, nb = observed.shape
    stride = ((nb + 1) // 2 * 3 + 3) // 4 * 4
    u8 = out if out is not None else np.zeros((ns, stride), np.uint8)
    assert u8.shape == (ns, stride) and u8.dtype == np.uint8 and u8.flags.c_contiguous
    idx, val = _pack_native(12, observed, u8, stride)
    return u8, idx, val


def pack_counts_numpy(observed):
    """pack_counts, written out in numpy (the definition the encoder is tested against)."""
    observed = np.asarray(observed)
    if observed.min(initial=0) < 0:
        raise ValueError("negative read count")
    big = observed >= 65535
    idx = np.flatnonzero(big.ravel()).astype(np.int64)
    val = observed.ravel()[idx].astype(np.int32)
    return np.minimum(observed, 65535).astype(np.uint16), idx, val


def pack_counts12_numpy(observed):
    """pack_counts12, written out in numpy (the definition the encoder is tested against)."""
    observed = np.asarray(observed)
    if observed.min(initial=0) < 0:
        raise ValueError("negative read count")
    ns, nb = observed.shape
    big = observed >= 4095
    idx = np.flatnonzero(big.ravel()).astype(np.int64)
    val = observed.ravel()[idx].astype(np.int32)
    pairs = (nb + 1) // 2
    u8 = np.zeros((ns, (pairs * 3 + 3) // 4 * 4), np.uint8)
    v = np.zeros((ns, 2 * pairs), np.uint16)
    np.minimum(observed, 4095, out=v[:, :nb], casting="unsafe")
    v0, v1 = v[:, 0::2], v[:, 1::2]
    u8[:, 0:3 * pairs:3] = v0 & 0xFF
    u8[:, 1:3 * pairs:3] = (v0 >> 8) | ((v1 & 0xF) << 4)
    u8[:, 2:3 * pairs:3] = v1 >> 4
    return u8, idx, val


class DeviceGraph:
    """A captured device-resident batch (Cohort.capture_device).  Holds the tensors it was captured over."""

    def __init__(self, lib, handle, tensors, status):
        self.lib, self.handle, self._tensors, self.status = lib, handle, tensors, status

    def launch(self, stream=None):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        return _lib.check(self.lib.edb200_graph_launch(self.handle, st), "edb200_graph_launch")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.edb200_graph_destroy(self.handle)
            self.handle = None
            self._tensors = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cohort:
    def __init__(self, chain_offsets, start, end, n_states=3, mixture=1.0, odds=None, transitions=None,
                 transition_probability=1e-4, expected_cnv_length=50000.0, build_table=True, device=-1):
        _lib.init(device)
        self.lib = _lib.load()
        self.chain_offsets = np.ascontiguousarray(np.asarray(chain_offsets, np.int64))
        self.start = np.ascontiguousarray(np.asarray(start, np.int32))
        self.end = np.ascontiguousarray(np.asarray(end, np.int32))
        self.n_bins = int(self.start.size)
        self.n_chains = int(self.chain_offsets.size - 1)
        self.n_states = int(n_states)
        odds_a = None if odds is None else np.ascontiguousarray(np.asarray(odds, np.float64))
        T_a = None if transitions is None else np.ascontiguousarray(np.asarray(transitions, np.float64).ravel(order="F"))
        spec = _lib.CohortSpec(self.n_bins, self.n_chains, _ptr(self.chain_offsets), _ptr(self.start), _ptr(self.end),
                               self.n_states, _ptr(odds_a), float(mixture), _ptr(T_a), float(transition_probability),
                               float(expected_cnv_length), 0 if build_table else 1)
        h = C.c_void_p()
        _lib.check(self.lib.edb200_cohort_create(C.byref(spec), C.byref(h)), "edb200_cohort_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.edb200_cohort_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Execution options (include/exomedepth_b200.h, EDB200_OPT_*): 'sweep' (0 auto, 1 lane per state, 2 thread per
        chain), 'parts', 'vsplit', 'crit_warps', 'sweep_warps', 'packplan', 'segments' (-1 auto, 0, 1), 'seg_warm', 'seg_min',
        'seg_repair'.  Results never depend on them."""
        _lib.check(self.lib.edb200_cohort_set_option(self.handle, _lib.OPTIONS[name], int(value)), "edb200_cohort_set_option")
        return self

    def segment_stats(self):
        """dict(pieces, listed, chains_repaired, pairs_repaired) of the last Viterbi pass if it was a segmented sweep
        (pieces = 0 otherwise); see include/exomedepth_b200.h: edb200_cohort_segment_stats."""
        out = (C.c_int32 * 10)()
        _lib.check(self.lib.edb200_cohort_segment_stats(self.handle, out), "edb200_cohort_segment_stats")
        why = dict(zip(("non_finite", "list_full", "seam_values", "seam_error", "on_path", "forced"), list(out)[4:10]))
        return dict(pieces=out[0], listed=out[1], chains_repaired=out[2], pairs_repaired=out[3], why=why)

    # ---- shared metadata ------------------------------------------------------------------------
    def table_bytes(self):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.edb200_cohort_table(self.handle, C.byref(p), C.byref(n)), "edb200_cohort_table")
        return int(n.value)

    def table_to(self, tensor, stream=0):
        """copy the log-transition table into a torch.float64 CUDA tensor (for dist.broadcast from rank 0)"""
        assert tensor.is_cuda and tensor.numel() * tensor.element_size() >= self.table_bytes()
        _lib.check(self.lib.edb200_cohort_table_copy(self.handle, tensor.data_ptr(), 0, stream), "table_copy")

    def table_from(self, tensor, stream=0):
        assert tensor.is_cuda and tensor.numel() * tensor.element_size() >= self.table_bytes()
        _lib.check(self.lib.edb200_cohort_table_copy(self.handle, tensor.data_ptr(), 1, stream), "table_copy")

    # ---- host buffers ---------------------------------------------------------------------------
    def run_host(self, observed, reference, phi, expected, want_ll=True, want_path=True, call_cap=512,
                 mode=_lib.EMISSION_AUTO, out=None, want_stats=False, overflow=None):
        """observed int32[n_samples, n_bins] — or uint16 / uint8 in the ingestion layouts of pack_counts / pack_counts12, with overflow = (flat indices,
        values) of the counts that do not fit 16 bits; reference int32[n_bins] (shared) or [n_samples, n_bins];
        phi, expected float64[n_samples].  Returns dict(ll [n,S,bins], path int8 [n,bins], calls, ncalls, status);
        with want_stats also call_stats float64[n, call_cap, 3] (per call: sum of ll[,type] - ll[,normal],
        sum of total*expected, sum of test; R/class_definition.R:393-400) and cor float64[n] (:338), both computed
        on the device — the likelihood matrix then only crosses PCIe if want_ll is set."""
        u16 = np.asarray(observed).dtype == np.uint16
        p12 = np.asarray(observed).dtype == np.uint8          # rows of 12-bit fields (pack_counts12)
        observed = np.asarray(observed, np.uint8 if p12 else np.uint16 if u16 else np.int32)
        if not (p12 and observed.ndim == 2 and observed.strides[1] == 1 and observed.strides[0] % 4 == 0):   # (12-bit rows may sit further apart than they are long)
            observed = np.ascontiguousarray(observed)
        reference = np.ascontiguousarray(np.asarray(reference, np.int32))
        ns = observed.shape[0]
        ovf_i = ovf_v = None
        if (u16 or p12) and overflow is not None and len(overflow[0]):
            ovf_i = np.ascontiguousarray(np.asarray(overflow[0], np.int64))
            ovf_v = np.ascontiguousarray(np.asarray(overflow[1], np.int32))
        S, nb = self.n_states, self.n_bins
        per_bin = np.ndim(phi) == 2 or np.ndim(expected) == 2     # [n_samples, n_bins]: per-bin fits (phi.bins > 1, covariates)
        shape = (ns, nb) if per_bin else (ns,)
        phi = np.ascontiguousarray(np.broadcast_to(np.asarray(phi, np.float64), shape))
        expected = np.ascontiguousarray(np.broadcast_to(np.asarray(expected, np.float64), shape))
        assert observed.shape == ((ns, ((nb + 1) // 2 * 3 + 3) // 4 * 4) if p12 else (ns, nb))
        out = out or {}
        ll = out.get("ll") if want_ll else None
        if want_ll and ll is None:
            ll = np.empty((ns, S, nb))
        path = out.get("path") if want_path else None
        if want_path and path is None:
            path = np.empty((ns, nb), np.int8)
        calls = out.get("calls")
        if calls is None:
            calls = np.zeros((ns, call_cap, 4), np.int32)
        ncalls = out.get("ncalls")
        if ncalls is None:
            ncalls = np.zeros(ns, np.int32)
        stats = cor = None
        if want_stats:
            stats = out.get("call_stats")
            if stats is None:
                stats = np.zeros((ns, call_cap, 3))
            cor = out.get("cor")
            if cor is None:
                cor = np.zeros(ns)
        b = _lib.Batch(ns, None if (u16 or p12) else _ptr(observed), nb, _ptr(reference), 0 if reference.ndim == 1 else nb, _ptr(phi),
                       _ptr(expected), _ptr(ll), nb, _ptr(path), nb, _ptr(calls), _ptr(ncalls), call_cap,
                       _ptr(stats), _ptr(cor), nb if per_bin else 0, _ptr(observed) if u16 else None, nb, 0 if ovf_i is None else ovf_i.size,
                       _ptr(ovf_i), _ptr(ovf_v), _ptr(observed) if p12 else None, observed.strides[0] if p12 else 0)
        rc = _lib.check(self.lib.edb200_cohort_run_host(self.handle, C.byref(b), mode), "edb200_cohort_run_host")
        self._last_ns = ns
        return dict(ll=ll, path=path, calls=calls, ncalls=ncalls, status=rc, call_stats=stats, cor=cor)

    def fit(self, observed, reference):
        """Beta-binomial fit of every sample against its reference (default formula of new('ExomeDepth'),
        R/class_definition.R:118-119, 168; betabin.py).  Returns (phi, expected) per sample for run_host / call_cnvs."""
        from . import betabin
        r = betabin.fit(observed, reference)
        bad = np.flatnonzero((r["info"] == -1) | (r["info"] == -2))
        if bad.size:
            raise _lib.EDB200Error(f"beta-binomial fit failed for sample {int(bad[0])}: {betabin.INFO[int(r['info'][bad[0]])]}")
        return r["phi"], r["expected"]

    def call_cnvs(self, observed, reference, phi, expected, chromosome_names=None, call_cap=512,
                  mode=_lib.EMISSION_AUTO, want_ll=False, want_path=False):
        """`new('ExomeDepth')`'s likelihood step + `CallCNVs` for every sample of the cohort
        (R/class_definition.R:184-189, 311-419): returns dict(CNV_calls = one list of rows per sample with the columns
        of x@CNV.calls, cor = x@cor.test.reference per sample, ll / path when asked for).  The per-call sums are
        taken on the device (callcnvs.cu); only signif(), as.integer() and the id strings are done here."""
        from .api import _signif
        import math
        res = self.run_host(observed, reference, phi, expected, want_ll=want_ll, want_path=want_path,
                            call_cap=call_cap, mode=mode, want_stats=True)
        names = chromosome_names if chromosome_names is not None else [str(c + 1) for c in range(self.n_chains)]
        log10e = math.log10(math.e)
        out = []
        for s in range(res["ncalls"].size):
            n = int(res["ncalls"][s])
            if n > call_cap:
                raise _lib.EDB200Error(f"sample {s}: {n} calls exceed call_cap={call_cap}")
            rows = []
            for k in range(n):
                sp, ep, typ, nex = (int(v) for v in res["calls"][s, k])
                bf, rexp, robs = (float(v) for v in res["call_stats"][s, k])
                chrom = str(names[int(np.searchsorted(self.chain_offsets, sp - 1, side="right")) - 1])
                st, en = float(self.start[sp - 1]), float(self.end[ep - 1])
                rexp_i = int(rexp) if math.isfinite(rexp) else 0                       # as.integer()
                rows.append(dict(start_p=sp, end_p=ep, type=typ, nexons=nex, start=st, end=en, chromosome=chrom,
                                 id=f"chr{chrom}:{int(st)}-{int(en)}".replace("chrchr", "chr"),
                                 BF=_signif(log10e * bf, 3), reads_expected=rexp_i, reads_observed=robs,
                                 reads_ratio=_signif(robs / rexp_i, 3) if rexp_i else float("inf")))
            out.append(rows)
        return dict(CNV_calls=out, cor=res["cor"], ll=res["ll"], path=res["path"], status=res["status"],
                    calls=res["calls"], ncalls=res["ncalls"], call_stats=res["call_stats"])

    def forward_last(self, tp_grid=None):
        """Forward log-likelihood per sample and grid point over the likelihoods of the most recent run_host call
        (still resident in HBM), and the index of the maximiser — the per-sample transition-probability MLE over the
        grid.  tp_grid=None: one column for the cohort's own transition matrix.  Extension (no reference counterpart)."""
        if tp_grid is None:
            grid, ng = None, 1
        else:
            grid = np.ascontiguousarray(np.asarray(tp_grid, np.float64))
            ng = grid.size
        ns = self._last_ns
        loglik = np.empty((ns, ng))
        best = np.empty(ns, np.int32)
        _lib.check(self.lib.edb200_cohort_forward_last(self.handle, _ptr(grid), ng, _ptr(loglik), _ptr(best)),
                   "edb200_cohort_forward_last")
        return loglik, best

    # ---- device tensors (torch) -----------------------------------------------------------------
    def _device_batch(self, observed, reference, phi, expected, ll, path, calls, ncalls, call_stats, cor):
        return _lib.Batch(observed.shape[0], observed.data_ptr(), observed.stride(0), reference.data_ptr(),
                          0 if reference.dim() == 1 else reference.stride(0), phi.data_ptr(), expected.data_ptr(),
                          ll.data_ptr(), ll.stride(1), path.data_ptr() if path is not None else None,
                          path.stride(0) if path is not None else 0, calls.data_ptr() if calls is not None else None,
                          ncalls.data_ptr() if ncalls is not None else None, calls.shape[1] if calls is not None else 0,
                          call_stats.data_ptr() if call_stats is not None else None, cor.data_ptr() if cor is not None else None)

    def run_device(self, observed, reference, phi, expected, ll, path=None, calls=None, ncalls=None,
                   what=3, mode=_lib.EMISSION_AUTO, stream=None, call_stats=None, cor=None):
        """All arguments are CUDA tensors on this cohort's device; enqueues on `stream` (default: torch's
        current stream) without synchronising."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        b = self._device_batch(observed, reference, phi, expected, ll, path, calls, ncalls, call_stats, cor)
        return _lib.check(self.lib.edb200_cohort_run_device(self.handle, C.byref(b), what, mode, st),
                          "edb200_cohort_run_device")

    def capture_device(self, observed, reference, phi, expected, ll, path=None, calls=None, ncalls=None,
                       what=3, mode=_lib.EMISSION_AUTO, call_stats=None, cor=None):
        """Same arguments as run_device: runs the batch once, records it into a CUDA graph and returns a
        DeviceGraph whose launch() replays it over the same tensors (their contents may change in between) —
        for small panels, where the step is bound by its launches (edb200_cohort_capture_device)."""
        b = self._device_batch(observed, reference, phi, expected, ll, path, calls, ncalls, call_stats, cor)
        h = C.c_void_p()
        rc = _lib.check(self.lib.edb200_cohort_capture_device(self.handle, C.byref(b), what, mode, C.byref(h)),
                        "edb200_cohort_capture_device")
        keep = (observed, reference, phi, expected, ll, path, calls, ncalls, call_stats, cor)
        return DeviceGraph(self.lib, h, keep, rc)

    def forward_device(self, ll, loglik, best=None, tp_grid=None, stream=None):
        """ll: CUDA float64 [n_samples, S, stride] as filled by run_device; loglik: CUDA float64 [n_samples, n_grid];
        best: CUDA int32 [n_samples] or None.  Enqueues on `stream` (default: torch's current stream)."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        grid = None if tp_grid is None else np.ascontiguousarray(np.asarray(tp_grid, np.float64))
        b = _lib.Batch(ll.shape[0], None, 0, None, 0, None, None, ll.data_ptr(), ll.stride(1), None, 0, None, None, 0)
        return _lib.check(self.lib.edb200_cohort_forward_device(self.handle, C.byref(b), _ptr(grid), 1 if grid is None else grid.size,
                                                                loglik.data_ptr(), best.data_ptr() if best is not None else None, st),
                          "edb200_cohort_forward_device")
