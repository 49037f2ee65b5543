"""Host-side mirror of `select.reference.set` (R/optimize_reference_set.R:51-148): bin selection on the host (a few
quantiles), the correlation sweep and the per-prefix beta-binomial fits on the GPU through the C ABI.

The reference fits with `aod::betabin` and scores with `VGAM::dbetabinom.ab` (third-party, unpinned): here the fit is
the likelihood maximiser of exomedepth_b200/csrc/betabin.cu and the score is restated from lgamma (betabin.py)."""
import ctypes as C

import numpy as np

from . import _lib


def select_bins(total_counts, bin_length=None, n_bins_reduced=0):
    """R/optimize_reference_set.R:81-88: the bins used to rank the candidate references (0-based indices).
    total_counts = test + all candidates; quantile() is R's default (type 7)."""
    total = np.asarray(total_counts, np.float64)
    bl = np.ones(total.size) if bin_length is None else np.asarray(bin_length, np.float64)
    if np.any(bl == 0):
        n0 = int(np.sum(bl == 0))
        raise ValueError(f"bin.length contains {n0} zero{'s' if n0 > 1 else ''}. This causes NAs in correlation computing. "
                         "All bin lengths must be positive")                                    # :69-72
    big = total > 30
    hi = np.quantile(total[big], 0.9) if np.any(big) else np.nan
    keep = big & (bl >= np.quantile(bl, 0.05)) & (bl <= np.quantile(bl, 0.95)) & (total < hi)
    sel = np.flatnonzero(keep)
    if 0 < n_bins_reduced < sel.size:                       # :88  selected[seq(1, length, length / n.bins.reduced)]
        step = sel.size / n_bins_reduced
        n = int(np.floor((sel.size - 1) / step + 1e-10)) + 1
        sel = sel[np.floor(1 + step * np.arange(n)).astype(np.int64) - 1]
    return sel.astype(np.int32)


def correlations(counts, selected, bin_length=None, row0=0, n_rows=None):
    """Pearson correlations of the normalised count rows over the selected bins (R/optimize_reference_set.R:100),
    rows row0 .. row0 + n_rows - 1 against every sample.  counts: int32[n_samples, n_bins]."""
    counts = np.ascontiguousarray(np.asarray(counts, np.int32))
    ns = counts.shape[0]
    n_rows = ns - row0 if n_rows is None else n_rows
    sel = np.ascontiguousarray(np.asarray(selected, np.int32))
    bl = None if bin_length is None else np.ascontiguousarray(np.asarray(bin_length, np.float64))
    out = np.empty((n_rows, ns))
    rc = _lib.load().edb200_refset_correlations(counts.ctypes.data, counts.shape[1], ns, None if bl is None else bl.ctypes.data,
                                                sel.ctypes.data, sel.size, row0, n_rows, out.ctypes.data)
    _lib.check(rc, "edb200_refset_correlations")
    return out


def cohort_reference_ranking(counts, bin_length=None, n_bins_reduced=0):
    """Leave-one-out sweep: every sample in turn is the test, all others the candidates.  Returns dict(selected,
    correlations [n, n], order [n, n-1] = for each test sample the other samples by decreasing correlation (:101))."""
    counts = np.asarray(counts)
    sel = select_bins(counts.sum(0, dtype=np.int64), bin_length, n_bins_reduced)
    cor = correlations(counts, sel, bin_length)
    n = cor.shape[0]
    order = np.empty((n, n - 1), np.int32)
    for t in range(n):
        others = np.array([i for i in range(n) if i != t])
        order[t] = others[np.argsort(-cor[t, others], kind="stable")]
    return dict(selected=sel, correlations=cor, order=order)


def rank_candidates(test_counts, reference_counts, bin_length=None, n_bins_reduced=0, names=None):
    """`select.reference.set` up to and including the ordering of the candidates (R/optimize_reference_set.R:51-102).
    reference_counts: bins x candidates, as in R.  Returns dict(ref_samples, correlations, selected, order) in the
    order of `summary.stats`."""
    test = np.asarray(test_counts)
    refs = np.asarray(reference_counts)
    if refs.ndim != 2:
        raise ValueError("The reference sequence count data must be provided as a matrix")     # :63
    if refs.shape[0] != test.size:
        raise ValueError("The number of rows of the reference matrix must match the length of the test count data")
    names = list(names) if names is not None else [f"X{i + 1}" for i in range(refs.shape[1])]  # :77
    if int(np.sum(test > 2)) < 5:                                                               # :55-60
        return dict(ref_samples=names[:1], correlations=None, selected=None)
    sel = select_bins(refs.sum(1, dtype=np.int64) + test, bin_length, n_bins_reduced)
    stacked = np.vstack([test[None, :], refs.T]).astype(np.int32)
    cor = correlations(stacked, sel, bin_length, row0=0, n_rows=1)[0, 1:]
    order = np.argsort(-cor, kind="stable")
    return dict(ref_samples=[names[i] for i in order], correlations=cor[order], selected=sel, order=order)


def select_reference_set(test_counts, reference_counts, bin_length=None, n_bins_reduced=0, names=None, chunk=32):
    """`select.reference.set(test.counts, reference.counts, bin.length, n.bins.reduced)` for the default formula and
    phi.bins = 1 (R/optimize_reference_set.R:51-148): candidates ordered by correlation, then the aggregate grown one
    candidate at a time — beta-binomial fit of the test against every prefix sum (GPU, `chunk` prefixes per call),
    expected Bayes factor of a heterozygous deletion (R/tools.R:128-166) — and the prefix with the largest expected
    BF chosen.  Returns dict(reference_choice, summary_stats) with summary_stats holding the columns of the R data
    frame (ref_samples, correlations, expected_BF, phi, RatioSd, mean_p, median_depth, selected) as arrays."""
    from . import betabin
    test = np.asarray(test_counts)
    refs = np.asarray(reference_counts)
    front = rank_candidates(test, refs, bin_length, n_bins_reduced, names)
    if front["correlations"] is None:
        return dict(reference_choice=front["ref_samples"], summary_stats=None)
    sel, order = front["selected"], front["order"]
    n = order.size
    t_sel = test[sel].astype(np.int32)
    r_sel = refs[sel][:, order].astype(np.int64)            # selected bins x candidates, best first
    cols = {k: np.full(n, np.nan) for k in ("expected_BF", "phi", "RatioSd", "mean_p", "median_depth")}
    running = np.zeros(sel.size, np.int64)
    done = False
    for i0 in range(0, n, chunk):
        i1 = min(i0 + chunk, n)
        prefix = running[None, :] + np.cumsum(r_sel[:, i0:i1].T, axis=0)      # :115  reference <- reference + reference.counts[, i]
        running = prefix[-1]
        fit = betabin.fit(np.tile(t_sel, (i1 - i0, 1)), prefix.astype(np.int32))
        # the expected Bayes factors of the whole chunk in one launch (R/optimize_reference_set.R:133-140); prefixes behind
        # the loop's break are simply not used
        med = np.median(prefix, axis=1)
        p_all = np.clip(fit["expected"], 1e-300, 1 - 1e-16)
        odds = p_all / (1 - p_all) * 0.5
        ok = (fit["info"] != -1) & (fit["info"] != -2) & (fit["phi"] > 0) & (fit["phi"] < 1)
        bf = np.full(i1 - i0, np.nan)
        if ok.any():
            bf[ok] = betabin.get_power_betabinom_batch(np.rint(med[ok]).astype(np.int32), fit["phi"][ok], p_all[ok], (odds / (1 + odds))[ok])
        for j in range(i1 - i0):
            i = i0 + j
            if fit["info"][j] == -1 or fit["info"][j] == -2:
                raise _lib.EDB200Error(f"beta-binomial fit of prefix {i + 1} failed: {betabin.INFO[int(fit['info'][j])]}")
            phi, p = float(fit["phi"][j]), float(fit["expected"][j])
            cols["phi"][i], cols["mean_p"][i] = phi, p                          # :125-126 (one phi, one p per fit)
            cols["median_depth"][i] = float(np.median(prefix[j]))               # :127
            cols["RatioSd"][i] = float(np.mean(np.sqrt(1 + (t_sel + prefix[j] - 1) * phi)))     # :128
            if i + 1 > 2 and p < 0.05:                                          # :130
                done = True
                break
            cols["expected_BF"][i] = bf[j]                                      # :133-140
        if done:
            break
    best = int(np.nanargmax(cols["expected_BF"]))                              # :143 which.max
    chosen = np.zeros(n, bool)
    chosen[best] = True
    stats = dict(ref_samples=front["ref_samples"], correlations=front["correlations"], selected=chosen, **cols)
    return dict(reference_choice=front["ref_samples"][:best + 1], summary_stats=stats)


# ---- device tensors (torch): the two stages, for the sharded sweep --------------------------------------------
def kpad(n_selected):
    return int(_lib.load().edb200_refset_kpad(int(n_selected)))


def standardize_device(counts, selected, bin_length, z, stream=None):
    """counts: CUDA int32 [n, stride]; selected: CUDA int32 [k]; bin_length: CUDA float64 [n_bins] or None;
    z: CUDA float64 [n, kpad(k)] (written), or the raw device address of such a block (block_alloc)."""
    import torch
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    return _lib.check(_lib.load().edb200_refset_standardize_device(
        counts.data_ptr(), counts.stride(0), counts.shape[0], bin_length.data_ptr() if bin_length is not None else None,
        selected.data_ptr(), selected.numel(), z if isinstance(z, int) else z.data_ptr(), st), "edb200_refset_standardize_device")


def block_alloc(rows_per_rank, n_selected):
    """This rank's block of standardised rows for the all-gather-free sharded sweep: (device address, 64-byte CUDA IPC
    handle to pass to the other ranks)."""
    p = C.c_void_p()
    h = C.create_string_buffer(64)
    _lib.check(_lib.load().edb200_refset_block_alloc(int(rows_per_rank), int(n_selected), C.byref(p), h), "edb200_refset_block_alloc")
    return int(p.value), h.raw


def peers_open(handles, my_rank):
    """handles: the 64-byte IPC handles of every rank's block, in rank order."""
    blob = b"".join(handles)
    _lib.check(_lib.load().edb200_refset_peers_open(blob, len(handles), int(my_rank)), "edb200_refset_peers_open")


def peers_close():
    _lib.check(_lib.load().edb200_refset_peers_close(), "edb200_refset_peers_close")


def gram_peers_device(m, rows_per_rank, n_total, n_selected, out, stream=None):
    """out[m, n_total] = this rank's first m block rows against every sample, the other ranks' rows read from their
    memory over NVLink (CUDA IPC) inside the Gram kernel."""
    import torch
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    return _lib.check(_lib.load().edb200_refset_gram_peers_device(int(m), int(rows_per_rank), int(n_total), int(n_selected),
                                                                  out.data_ptr(), st), "edb200_refset_gram_peers_device")


def gram_device(za, zb, n_selected, out, stream=None):
    """out[m, n] = correlations of the rows of za against the rows of zb."""
    import torch
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    return _lib.check(_lib.load().edb200_refset_gram_device(za.data_ptr(), za.shape[0], zb.data_ptr(), zb.shape[0], int(n_selected),
                                                            out.data_ptr(), st), "edb200_refset_gram_device")
