"""numpy restatement of the R side of the hot path: CallCNVs framing.  TEST INFRASTRUCTURE ONLY.

Follows R/class_definition.R:311-419 (CallCNVs) and R/tools.R:88-103 (viterbi.hmm): bin ordering,
transition matrix, per-chromosome dummy first/last observation, −1 index shift, BF / reads.* columns.
The emission and HMM kernels are passed in (compiled reference, C port, or the CUDA path under test).
"""
import math

import numpy as np


def signif(x, digits=3):
    """R's signif() for scalars (ties are measure-zero for the float sums used here)."""
    x = float(x)
    if x == 0 or not math.isfinite(x):
        return x
    e = math.floor(math.log10(abs(x)))
    return round(x, digits - 1 - e)


def _lsum(v):
    v = np.asarray(v, float)
    return float(math.fsum(v)) if np.all(np.isfinite(v)) else float(np.sum(v))


def chromosome_levels(chromosome):
    """R/class_definition.R:323-326 — '1'..'22' first, then the others in order of appearance."""
    used = list(dict.fromkeys(str(c) for c in chromosome))
    auto = [str(i) for i in range(1, 23)]
    levels = auto + [c for c in used if c not in auto]
    return [c for c in levels if c in used]


def bin_order(chromosome, start, end):
    """R/class_definition.R:328-329 — order(chromosome factor, 0.5*(start+end)), stable."""
    levels = chromosome_levels(chromosome)
    code = np.array([levels.index(str(c)) for c in chromosome])
    mid = 0.5 * (np.asarray(start, float) + np.asarray(end, float))
    return np.lexsort((mid, code)), code, levels


def transition_matrix(tp, n_states=3):
    """R/class_definition.R:343-347 (S=3); SURVEY §8a H4 for S>3. T[k, j] = P(k -> j)."""
    T = np.zeros((n_states, n_states))
    T[0, 0] = 1.0 - tp
    T[0, 1:] = tp / (n_states - 1.0)
    for k in range(1, n_states):
        T[k, 0] = 0.5
        T[k, k] = 0.5
    return T


def hmm_column_order(n_states):
    """R/class_definition.R:364 `c(2, 1, 3)`: normal first, then the other states in copy-number order.
    For S=3 (CN 1,2,3) normal is column 1; for S=5/7 (CN 0..) normal is column 2."""
    normal = 1 if n_states == 3 else 2
    return [normal] + [s for s in range(n_states) if s != normal]


def frame_chromosome(ll_rows, start, end, L):
    """R/class_definition.R:364-368 — returns (loglik nobs×S in HMM order, positions int32[nobs])."""
    S = ll_rows.shape[1]
    cols = hmm_column_order(S)
    head = np.full((1, S), -np.inf)
    head[0, 0] = 0.0
    tail = np.full((1, S), -100.0)
    tail[0, 0] = 0.0
    loc = np.vstack([head, ll_rows[:, cols], tail])
    pos = np.concatenate([[start[0] - 2 * L], start, [end[-1] + 2 * L]])
    return loc, np.trunc(pos).astype(np.int32)   # as.integer()


def call_cnvs(likelihood, test, reference, expected, chromosome, start, end, hmm,
              transition_probability=1e-4, expected_cnv_length=50000):
    """CallCNVs (R/class_definition.R:311-419) on top of `hmm(T, loglik, positions, L) -> (path, calls)`.

    Returns dict(order, cor, paths {chrom: path}, calls [list of dict rows]).
    """
    likelihood = np.asarray(likelihood, float)
    test = np.asarray(test, float)
    reference = np.asarray(reference, float)
    expected = np.asarray(expected, float)
    start = np.asarray(start, float)
    end = np.asarray(end, float)
    S = likelihood.shape[1]
    order, code, levels = bin_order(chromosome, start, end)
    chrom = np.array([str(c) for c in chromosome], dtype=object)
    if np.any(order != np.arange(order.size)):
        # :331-336 — test, reference, annotations, likelihood are reordered; expected is NOT
        test, reference, likelihood = test[order], reference[order], likelihood[order]
        chrom, code, start, end = chrom[order], code[order], start[order], end[order]
    cor = float(np.corrcoef(test, reference)[0, 1]) if test.size > 1 else float("nan")
    total = test + reference
    T = transition_matrix(transition_probability, S)
    cols = hmm_column_order(S)
    rows, paths = [], {}
    shift = 0
    for c in dict.fromkeys(code.tolist()):
        good = np.nonzero(code == c)[0]
        loc_ll = likelihood[good]
        loc, pos = frame_chromosome(loc_ll, start[good], end[good], expected_cnv_length)
        path, calls = hmm(T, loc, pos, float(expected_cnv_length))
        paths[levels[c]] = path
        for (sp, ep, typ, nex) in calls:
            sp0, ep0 = int(sp) - 1, int(ep) - 1                     # :371-372 (1-based into `good`)
            sl = slice(sp0 - 1, ep0) if sp0 >= 1 else slice(0, 0)
            col_type = cols[int(typ)]
            # R's sum() accumulates in long double (:395-400); math.fsum (exactly rounded) is the nearest stand-in
            bf = _lsum(loc_ll[sl, col_type] - loc_ll[sl, cols[0]])
            rexp = _lsum(total[good][sl] * expected[good][sl])
            robs = _lsum(test[good][sl])
            rexp_i = int(rexp) if math.isfinite(rexp) else 0
            rows.append(dict(
                start_p=sp0 + shift, end_p=ep0 + shift, type=int(typ), nexons=int(nex),
                start=float(start[good][sp0 - 1]) if sp0 >= 1 else float("nan"),
                end=float(end[good][ep0 - 1]) if ep0 >= 1 else float("nan"),
                chromosome=levels[c],
                BF=signif(math.log10(math.e) * bf, 3), BF_raw=bf, reads_expected_raw=rexp,
                reads_expected=rexp_i, reads_observed=robs,
                reads_ratio=signif(robs / rexp_i, 3) if rexp_i else float("inf"),
            ))
        shift += good.size
    return dict(order=order, cor=cor, paths=paths, calls=rows, transitions=T)
