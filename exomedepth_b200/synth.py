"""Synthetic cohorts of SURVEY.md §8d: exome-like bin geometry, a shared reference aggregate, per-sample
beta-binomial test counts with planted CNV segments.  numpy only; used by bench.py and the tests."""
import os

import numpy as np

SEED = 20261017
_GEOM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                     "exons_hg19_geometry.npz")


def geometry(n_bins=200_000, per_chrom=None):
    """(chain_offsets int64[n_chr+1], start int32[n], end int32[n]).

    Bins are the hg19 exons of the reference's data/bedFiles/exons_hg19.bed (committed as
    tests/golden/exons_hg19_geometry.npz; start = BED start + 1 as R/countBamInGranges.R:326), ordered by
    (chromosome, midpoint) as CallCNVs orders them, padded to `n_bins` with one synthetic chromosome — or,
    with per_chrom=(n_chr, bins), the first `bins` bins of the first n_chr chromosomes (small-panel config)."""
    g = np.load(_GEOM)
    code, start, end = g["chrom_code"].astype(np.int64), g["start"].astype(np.int64) + 1, g["end"].astype(np.int64)
    order = np.lexsort((0.5 * (start + end), code))
    code, start, end = code[order], start[order], end[order]
    if per_chrom is not None:
        n_chr, nb = per_chrom
        keep = np.concatenate([np.nonzero(code == c)[0][:nb] for c in range(n_chr)])
        code, start, end = code[keep], start[keep], end[keep]
    elif n_bins <= code.size:
        # keep whole-genome proportions: take the first n_bins*share bins of every chromosome
        share = n_bins / code.size
        keep = []
        for c in np.unique(code):
            idx = np.nonzero(code == c)[0]
            keep.append(idx[:max(2, int(round(idx.size * share)))])
        keep = np.concatenate(keep)[:n_bins]
        code, start, end = code[keep], start[keep], end[keep]
    else:
        extra = n_bins - code.size
        rng = np.random.default_rng(SEED - 1)
        gaps = rng.integers(200, 20000, extra)
        s = 10_000 + np.cumsum(gaps)
        code = np.concatenate([code, np.full(extra, code.max() + 1)])
        start = np.concatenate([start, s])
        end = np.concatenate([end, s + rng.integers(50, 400, extra)])
    bounds = np.nonzero(np.diff(code))[0] + 1
    offsets = np.concatenate([[0], bounds, [code.size]]).astype(np.int64)
    return offsets, start.astype(np.int32), end.astype(np.int32)


def shared(n_bins):
    """per-bin base rate and the shared reference aggregate (int32[n_bins])"""
    rng = np.random.default_rng(SEED)
    lam = rng.gamma(1.2, 80.0, n_bins)
    lam[rng.random(n_bins) < 0.15] = 0.0
    ref = rng.poisson(10.0 * lam).astype(np.int32)
    return lam, ref


def sample(s, ref, n_segments=100):
    """(observed int32[n_bins], phi, expected) for sample index s"""
    rng = np.random.default_rng(SEED + 1 + s)
    n = ref.size
    e = rng.uniform(0.08, 0.30)
    phi = float(np.exp(rng.uniform(np.log(5e-4), np.log(1e-2))))
    a1, a2 = e * (1 - phi) / phi, (1 - e) * (1 - phi) / phi
    p = rng.beta(a1, a2, n)
    cn = np.ones(n)
    for q in range(n_segments):
        b0 = int(rng.integers(0, n))
        ln = int(rng.integers(1, 21))
        cn[b0:b0 + ln] = 0.5 if q % 2 == 0 else 1.5
    obs = rng.poisson(ref * cn * p / (1 - p)).astype(np.int32)
    return obs, phi, e


def cohort(n_samples, n_bins=200_000, per_chrom=None, first_sample=0):
    """dict(offsets, start, end, reference, observed [n_samples, n_bins], phi, expected)"""
    offsets, start, end = geometry(n_bins, per_chrom)
    nb = start.size
    _, ref = shared(nb)
    obs = np.empty((n_samples, nb), np.int32)
    phi = np.empty(n_samples)
    exp = np.empty(n_samples)
    for i in range(n_samples):
        obs[i], phi[i], exp[i] = sample(first_sample + i, ref)
    return dict(offsets=offsets, start=start, end=end, reference=ref, observed=obs, phi=phi, expected=exp)
