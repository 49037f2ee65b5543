"""numpy restatement of the deterministic front half of select.reference.set.  TEST INFRASTRUCTURE ONLY.

Follows R/optimize_reference_set.R:
  :81-88    bins used for the selection (total.counts > 30, bin.length within its 5 % / 95 % quantiles, total.counts
            below the 90 % quantile of the bins above 30), optional grid sub-sampling (:88)
  :100      my.correlations: cor(x / (bin.length * sum(x) / 1e6), test / (bin.length * sum(test) / 1e6)) per reference
  :101-102  references ordered by decreasing correlation
The greedy aggregate loop that follows (:113-141) re-fits the beta-binomial model per prefix with aod::betabin and
scores it with VGAM::dbetabinom.ab (third-party, unpinned): select_reference_set() below restates it on the scipy
stand-ins of oracle/betabin.py.

With leave-one-out cohorts (every sample in turn is the test, all others the candidates) total.counts and the bin
filter do not depend on which sample is the test, so the whole sweep is one N x N Pearson matrix.

Parity unpinned: R is not available here and the reference has no fixture for this function; quantile() is R's default
type 7 (numpy's "linear"), cor() is the textbook Pearson coefficient.
"""
import numpy as np


def select_bins(total_counts, bin_length=None, n_bins_reduced=0):
    """R/optimize_reference_set.R:81-88 -> 0-based indices of the selected bins."""
    total = np.asarray(total_counts, float)
    bl = np.ones(total.size) if bin_length is None else np.asarray(bin_length, float)
    if np.any(bl == 0):
        raise ValueError("bin.length contains zeros. All bin lengths must be positive")       # :69-72
    above = total[total > 30]
    q = np.quantile(above, [0.1, 0.9]) if above.size else np.array([np.nan, np.nan])
    sel = np.nonzero((total > 30) & (bl >= np.quantile(bl, 0.05)) & (bl <= np.quantile(bl, 0.95)) & (total < q[1]))[0]
    if 0 < n_bins_reduced < sel.size:
        # selected[seq(1, length(selected), length(selected) / n.bins.reduced)]: fractional indices truncate
        step = sel.size / n_bins_reduced
        grid = 1 + step * np.arange(int(np.floor((sel.size - 1) / step + 1e-10)) + 1)
        sel = sel[np.floor(grid).astype(np.int64) - 1]
    return sel


def normalised(x, bin_length):
    """x / (bin.length * sum(x) / 10^6) with R's left-to-right evaluation (:100)."""
    x = np.asarray(x, float)
    return x / (bin_length * x.sum() / 1e6)


def correlations(test, references, bin_length=None):
    """:100 for already-selected rows: references is bins x n_ref; returns n_ref correlations."""
    test = np.asarray(test, float)
    references = np.asarray(references, float)
    bl = np.ones(test.size) if bin_length is None else np.asarray(bin_length, float)
    t = normalised(test, bl)
    return np.array([np.corrcoef(normalised(references[:, j], bl), t)[0, 1] for j in range(references.shape[1])])


def cohort_correlations(counts, bin_length=None, n_bins_reduced=0):
    """Leave-one-out sweep over a cohort: counts is n_samples x n_bins.
    Returns (selected bin indices, n_samples x n_samples correlation matrix over those bins)."""
    counts = np.asarray(counts, float)
    sel = select_bins(counts.sum(0), bin_length, n_bins_reduced)
    bl = np.ones(counts.shape[1]) if bin_length is None else np.asarray(bin_length, float)
    y = np.stack([normalised(counts[s, sel], bl[sel]) for s in range(counts.shape[0])])
    return sel, np.corrcoef(y)


def ranking(cor_row, self_index):
    """:101 order(my.correlations, decreasing = TRUE) over the other samples (stable, like R's order)."""
    idx = np.array([i for i in range(cor_row.size) if i != self_index])
    return idx[np.argsort(-cor_row[idx], kind="stable")]


def select_reference_set(test, references, bin_length=None, n_bins_reduced=0, names=None):
    """The whole function for the default formula / phi.bins = 1 (R/optimize_reference_set.R:51-148), with the scipy
    stand-ins of oracle/betabin.py for aod::betabin and VGAM::dbetabinom.ab (parity unpinned)."""
    from oracle import betabin as obb
    test = np.asarray(test, float)
    references = np.asarray(references, float)
    names = list(names) if names is not None else [f"X{i + 1}" for i in range(references.shape[1])]
    sel = select_bins(references.sum(1) + test, bin_length, n_bins_reduced)
    bl = np.ones(test.size) if bin_length is None else np.asarray(bin_length, float)
    cor = correlations(test[sel], references[sel], bl[sel])
    order = np.argsort(-cor, kind="stable")
    t, r = test[sel], references[sel][:, order]
    n = order.size
    cols = {k: np.full(n, np.nan) for k in ("expected_BF", "phi", "RatioSd", "mean_p", "median_depth")}
    reference = np.zeros(sel.size)
    for i in range(n):
        reference = reference + r[:, i]
        mu, phi, _ = obb.fit(t, reference)
        cols["phi"][i], cols["mean_p"][i] = phi, mu
        cols["median_depth"][i] = np.median(reference)
        cols["RatioSd"][i] = np.mean(np.sqrt(1 + (t + reference - 1) * phi))
        if i + 1 > 2 and mu < 0.05:
            break
        alt_odds = mu / (1 - mu) * 0.5
        cols["expected_BF"][i] = obb.get_power_betabinom(round(cols["median_depth"][i]), phi, mu, alt_odds / (1 + alt_odds))
    best = int(np.nanargmax(cols["expected_BF"]))
    return dict(reference_choice=[names[j] for j in order[:best + 1]], ref_samples=[names[j] for j in order],
                correlations=cor[order], **cols)
