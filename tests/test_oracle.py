"""Pin the oracle: the C restatement (oracle/oracle.c) against (a) the known-answer vectors of
SURVEY.md §8c, (b) committed outputs of the reference's own compiled C/C++ (tests/golden/ref_vectors.npz,
made by tools/make_golden.py) and (c) the compiled reference itself when its .so is present."""
import numpy as np
import pytest

from conftest import hmm_cases
from oracle import framing


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


# ---------------------------------------------------------------- KAT-1..3
def test_kat1_viterbi_doc_example(port, kat):
    k = kat["kat1"]
    path, calls = port.c_hmm(k["T"], k["loglik"], k["positions"], k["L"])
    assert path.tolist() == k["path"]
    assert calls.tolist() == k["calls"]          # includes the stale start 4 of hmm.cpp:111-121
    path, calls = port.c_hmm(np.eye(3), k["loglik"], k["positions"], k["L"])   # R/tools.R:82-85
    assert path.tolist() == [0] * 10 and calls.shape == (0, 4)


def test_kat2_lnbeta(port, kat):
    for x, y, v in kat["kat2"]:
        assert abs(float(port.lnbeta(x, y)) - v) <= 4e-15 * abs(v)


def test_kat3_loglike(port, kat):
    for tot, obs, mix, *vals in kat["kat3"]:
        got = port.get_loglike_matrix([kat["kat3_phi"]], [kat["kat3_expected"]], [tot], [obs], mix)[0]
        np.testing.assert_allclose(got, vals, rtol=1e-14, atol=0)


# ---------------------------------------------------------------- committed reference outputs
def test_lnbeta_matches_committed_reference(port, refvec):
    got = port.lnbeta(refvec["lnbeta_x"], refvec["lnbeta_y"])
    want = refvec["lnbeta_val"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    # bit-identical everywhere except the near-negative-integer series, which the port restates with
    # closed-form polygamma values (VP_gamma.c:795-894 reaches them through psi/zeta tables)
    exact = got[ok] == want[ok]
    assert exact.mean() > 0.97
    np.testing.assert_allclose(got[ok], want[ok], rtol=5e-12, atol=0)
    pos = ok & (refvec["lnbeta_x"] > 0) & (refvec["lnbeta_y"] > 0)
    assert np.array_equal(got[pos], want[pos])


def test_emission_matches_committed_reference(port, refvec):
    args = [refvec[k] for k in ("em_phi", "em_expected", "em_total", "em_observed")]
    assert same(port.get_loglike_matrix(*args, 1.0), refvec["em_ll_mix1"])
    assert same(port.get_loglike_matrix(*args, 0.4), refvec["em_ll_mix04"])
    assert np.isnan(refvec["em_ll_mix1"]).sum() > 100      # the pathological-phi rows are exercised
    zero = refvec["em_total"] == 0
    assert np.all(refvec["em_ll_mix1"][zero][~np.isnan(refvec["em_ll_mix1"][zero])] == 0.0)


def test_hmm_matches_committed_reference(port, refvec):
    n_nan_cases = 0
    for T, ll, pos, L, path, calls in hmm_cases(refvec):
        p, c = port.c_hmm(T, ll, pos, L)
        assert np.array_equal(p, path)
        assert np.array_equal(c, calls)
        n_nan_cases += int(np.isnan(port.log_transition_table(T, pos, L)).any())
    assert n_nan_cases >= 5     # negative distances -> log(negative) = NaN edge


def test_kat4_exomecount_end_to_end(port, refvec, exomecount, kat):
    ec = exomecount
    test = ec["Exome4"].astype(float)
    reference = (ec["Exome1"] + ec["Exome2"] + ec["Exome3"]).astype(float)
    n = test.size
    e = np.full(n, kat["kat3_expected"])
    ll = port.get_loglike_matrix(np.full(n, kat["kat3_phi"]), e, (test + reference).astype(np.int32),
                                 test.astype(np.int32), 1.0)
    assert np.array_equal(ll, refvec["kat4_ll"])
    np.testing.assert_allclose(ll.sum(0), kat["kat4"]["colsums"], rtol=1e-13)
    res = framing.call_cnvs(ll, test, reference, e, ["chr1"] * n, ec["start"], ec["end"], port.c_hmm)
    assert np.array_equal(res["paths"]["chr1"], refvec["kat4_path"])
    calls = res["calls"]
    k4 = kat["kat4"]
    assert len(calls) == k4["ncalls"]
    assert sum(c["type"] == 1 for c in calls) == k4["ndel"] and sum(c["type"] == 2 for c in calls) == k4["ndup"]
    assert sum(c["nexons"] for c in calls) == k4["nbins_cnv"]
    first = calls[0]
    assert (first["start_p"], first["end_p"], first["nexons"], first["BF"], first["reads_observed"],
            first["reads_expected"]) == (25, 27, 3, 12.4, 68, 224)
    keys = ("start_p", "end_p", "type", "nexons", "start", "end", "BF", "reads_expected", "reads_observed", "reads_ratio")
    got = np.array([[c[k] for k in keys] for c in calls], float)
    assert np.array_equal(got, refvec["kat4_calls"])


# ---------------------------------------------------------------- live compiled reference (dev container)
def test_port_vs_live_reference_dense(port, ref):
    rng = np.random.default_rng(7)
    x = 10 ** rng.uniform(-3, 5, 20000)
    y = 10 ** rng.uniform(-3, 5, 20000)
    assert np.array_equal(port.lnbeta(x, y), ref.lnbeta(x, y))
    g = 10 ** rng.uniform(-2, 6, 5000)
    assert np.array_equal(port.gammastar(g), ref.gammastar(g))
    u = rng.uniform(-0.9, 3, 5000) * rng.choice([1, 1e-2, 1e-4], 5000)
    assert np.array_equal(port.log1plusx(u), ref.log1plusx(u))
    n = 20000
    phi = 10 ** rng.uniform(-3.5, -0.01, n)
    e = rng.uniform(0.01, 0.9, n)
    tot = rng.poisson(10 ** rng.uniform(0, 4, n)).astype(np.int32)
    obs = rng.binomial(tot, e).astype(np.int32)
    for mix in (1.0, 0.3):
        assert same(port.get_loglike_matrix(phi, e, tot, obs, mix), ref.get_loglike_matrix(phi, e, tot, obs, mix))


def test_port_hmm_vs_live_reference(port, ref):
    rng = np.random.default_rng(8)
    for trial in range(100):
        nobs = int(rng.integers(2, 3000))
        ll = -rng.exponential(4, (nobs, 3))
        if trial % 2:
            ll = np.round(ll * 2) / 2
        pos = np.cumsum(rng.integers(-800, 20000, nobs)).astype(np.int32)
        tp = 10 ** rng.uniform(-6, -1)
        T = framing.transition_matrix(tp) if trial % 3 else rng.dirichlet(np.ones(3), 3)
        a, b = port.c_hmm(T, ll, pos, 50000.0), ref.c_hmm(T, ll, pos, 50000.0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_reference_rejects_other_state_counts(ref):
    """hmm.cpp:37-40 — prints an error and returns NULL for nstates != 3."""
    ll = np.zeros((5, 5))
    assert ref.c_hmm(np.full((5, 5), 0.2), ll, np.arange(5), 1.0) is None


# ---------------------------------------------------------------- extensions (unpinned): internal consistency
def test_sstate_definitions_reduce_to_reference_at_s3(port):
    assert np.array_equal(port.state_odds(3, 0.7), port.reference_odds(0.7))
    assert np.array_equal(port.callcnvs_transitions(3, 1e-4), framing.transition_matrix(1e-4))
    assert np.array_equal(port.callcnvs_transitions(5, 1e-4), framing.transition_matrix(1e-4, 5))
    o5 = port.state_odds(5)
    assert o5.tolist() == [0.05, 0.5, 1.0, 1.5, 2.0]


def test_forward_loglik_bounds_viterbi(port):
    """log-sum-exp over paths >= the best single path; equal when only one path is possible."""
    rng = np.random.default_rng(9)
    for S in (3, 5):
        nobs = 200
        ll = -rng.exponential(3, (nobs, S))
        ll[0] = [0] + [-np.inf] * (S - 1)
        pos = np.cumsum(rng.integers(1, 20000, nobs)).astype(np.int32)
        T = port.callcnvs_transitions(S, 1e-3)
        fw = port.forward_loglik(T, ll, pos, 50000.0)
        path, _ = port.c_hmm(T, ll, pos, 50000.0)
        lt = port.log_transition_table(T, pos, 50000.0)
        best = sum(ll[i, path[i]] + lt[i, path[i], path[i - 1]] for i in range(1, nobs))
        assert fw >= best - 1e-9
        assert fw < best + nobs * np.log(S)
    T = np.eye(3)
    ll = -rng.exponential(3, (50, 3))
    fw = port.forward_loglik(T, ll, np.arange(50, dtype=np.int32), 1.0)
    assert abs(fw - ll[1:, 0].sum()) < 1e-10


# ---------------------------------------------------------------- round-2 vectors (tools/make_golden_r2.py)
def test_round2_vectors_match_the_port(port, refvec2):
    """Leave-one-out ExomeCount, the parameter envelope and the a1 ~ -1 rows: the C restatement reproduces the compiled
    reference's likelihood matrices bit for bit (NaN cells included) and, through the CallCNVs framing, its paths and calls."""
    r = refvec2
    for pre in ("env", "negint"):
        got = port.get_loglike_matrix(r[f"{pre}_phi"], r[f"{pre}_expected"], r[f"{pre}_total"], r[f"{pre}_observed"], 1.0)
        assert same(got, r[f"{pre}_ll"]), pre
    assert np.isnan(r["negint_ll"]).all()
    ec = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "exomecount.npz"))
    names = ["Exome1", "Exome2", "Exome3", "Exome4"]
    n = ec["start"].size
    for s, nm in enumerate(names):
        test = ec[nm].astype(float)
        reference = sum(ec[o].astype(float) for o in names if o != nm)
        phi, e = float(r[f"loo{s}_phi"][0]), float(r[f"loo{s}_expected"][0])
        ll = port.get_loglike_matrix(np.full(n, phi), np.full(n, e), (test + reference).astype(np.int32), test.astype(np.int32), 1.0)
        assert same(ll, r[f"loo{s}_ll"])
        res = framing.call_cnvs(ll, test, reference, np.full(n, e), ["chr1"] * n, ec["start"], ec["end"], port.c_hmm)
        assert np.array_equal(res["paths"]["chr1"], r[f"loo{s}_path"])
        assert len(res["calls"]) == r[f"loo{s}_calls"].shape[0] > 20


def test_envelope_reference_error_vs_mpmath(refvec2):
    """Who is off where the GPU path and the reference disagree by more than 1e-10?  A 50-digit evaluation of
    ln B(a1+k, a2+n-k) - ln B(a1, a2) with the reference's own a1, a2: for small phi (a1 + a2 in the tens of thousands)
    the REFERENCE's lnbeta differences carry 1e-11 .. 1e-10 relative errors on bins with a handful of reads — its 1e-16
    roundings act on lnbeta terms of magnitude 1e5 that cancel down to a result of magnitude 1 (SURVEY.md §7 hard part 3);
    with thousands of reads the result is large and the relative error drops below 1e-13."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    r = refvec2
    phi, e, tot, obs, ll = r["env_phi"], r["env_expected"], r["env_total"], r["env_observed"], r["env_ll"]
    pick = np.flatnonzero((phi < 1e-4) & (tot > 0) & (tot <= 20) & np.isfinite(ll).all(1))[:80]
    assert pick.size >= 40
    worst = 0.0
    for i in pick:
        sd2 = phi[i] * e[i] * (1 - e[i])                       # src/CNV_estimate.cpp:73 (normal state: odds 1)
        a1 = e[i] * e[i] * (1 - e[i]) / sd2 - e[i]             # :45
        a2 = (1 - e[i]) / e[i] * a1                            # :46
        x, y = mp.mpf(a1) + int(obs[i]), mp.mpf(a2) + int(tot[i] - obs[i])
        exact = (mp.loggamma(x) + mp.loggamma(y) - mp.loggamma(x + y)) - (mp.loggamma(a1) + mp.loggamma(a2) - mp.loggamma(mp.mpf(a1) + mp.mpf(a2)))
        worst = max(worst, abs(float((mp.mpf(ll[i, 1]) - exact) / exact)))
    assert 1e-11 < worst < 1e-8, worst        # the reference itself is not a 1e-10 oracle out there
