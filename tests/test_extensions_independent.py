"""The S-state extensions (S = 5, 7 Viterbi; forward log-likelihood) have no reference code: oracle/oracle.c is their
definition ("parity unpinned").  This file holds a SECOND, independent restatement — pure Python, written from the
reference's hmm.cpp (src/hmm.cpp:42-126) with the S-state transition rule of SURVEY.md §8a H4, and a forward pass in
50-digit arithmetic (mpmath) — and checks that the oracle port agrees with it: bit-exactly for the integer outputs,
to 1e-12 for the forward log-likelihood.  CPU only."""
import math

import numpy as np
import pytest

INF = float("inf")


def log_or_nan(t):
    if t > 0:
        return math.log(t)
    return -INF if t == 0 else float("nan")


def viterbi_py(T, ll, pos, L):
    """T[k][j] = P(k -> j); ll[i][j] in HMM state order; returns (path, calls) like C_hmm (1-based calls)."""
    S, nobs = len(T), len(ll)
    V = [0.0] + [-INF] * (S - 1)
    frm = [[-1] * S]
    for i in range(1, nobs):
        d = math.exp(-(float(pos[i]) - float(pos[i - 1])) / L)
        nV, f = [-INF] * S, [-1] * S
        for j in range(S):
            for k in range(S):
                t = T[0][j] if k == 0 else d * T[k][j] + (1.0 - d) * T[0][j]
                newp = (ll[i][j] + V[k]) + log_or_nan(t)
                if newp > nV[j]:
                    nV[j], f[j] = newp, k
            if ll[i][j] == -INF:
                f[j] = 0
        V = nV
        frm.append(f)
    path = [0] * nobs
    for i in range(nobs - 1, 0, -1):
        st = path[i]
        path[i - 1] = 0 if st < 0 else frm[i][st]
    calls, current, start, nex = [], 0, -1, 0
    for i in range(1, nobs):
        if path[i - 1] != path[i]:
            if current == 0:
                start = i
            else:
                calls.append((start + 1, i, current, nex))
                nex = 0
        if path[i] != 0:
            nex += 1
        current = path[i]
    return path, calls


def forward_mp(T, ll, pos, L):
    import mpmath as mp
    mp.mp.dps = 50
    S, nobs = len(T), len(ll)
    a = [mp.mpf(1)] + [mp.mpf(0)] * (S - 1)            # linear domain, exact enough at 50 digits for these sizes
    scale = mp.mpf(0)
    for i in range(1, nobs):
        d = math.exp(-(float(pos[i]) - float(pos[i - 1])) / L)          # the model's own FP64 decay
        b = []
        for j in range(S):
            s = mp.mpf(0)
            for k in range(S):
                t = T[0][j] if k == 0 else d * T[k][j] + (1.0 - d) * T[0][j]
                if t > 0:                                                # t <= 0: log is -Inf / NaN, the term is skipped
                    s += a[k] * mp.mpf(t)
            b.append(s * mp.e ** mp.mpf(ll[i][j]) if ll[i][j] != -INF else mp.mpf(0))
        m = max(b)
        if m == 0:
            return -INF
        a = [x / m for x in b]
        scale += mp.log(m)
    return float(scale + mp.log(a[0])) if a[0] > 0 else -INF


def random_case(rng, S, nobs, rough=False):
    ll = -np.abs(rng.normal(0, 6, (nobs, S)))
    seg = 0
    while seg < nobs:                                    # planted runs favouring one state
        ln = int(rng.integers(1, 12))
        st = int(rng.integers(0, S))
        ll[seg:seg + ln, st] += 8.0
        seg += ln
    ll = np.minimum(ll, 0.0)
    if rough:
        ll[rng.integers(0, nobs, 3), rng.integers(1, S, 3)] = -INF
        ll[0] = [-INF if j else 0.0 for j in range(S)]                 # CallCNVs' dummy first row is never read
    pos = np.cumsum(rng.integers(50, 30000, nobs)).astype(np.int32)
    if rough:
        pos[nobs // 2] = pos[nobs // 2 - 1] - 700                      # a negative gap: log of a negative term (NaN)
        pos[nobs // 3] = pos[nobs // 3 - 1]                            # and a zero gap
    return ll, pos


@pytest.mark.parametrize("S", [3, 5, 7])
@pytest.mark.parametrize("rough", [False, True])
def test_port_viterbi_agrees_with_an_independent_restatement(port, S, rough):
    rng = np.random.default_rng(100 * S + rough)
    for case in range(6):
        nobs = int(rng.integers(2, 90))
        ll, pos = random_case(rng, S, nobs, rough)
        tp = float(10 ** rng.uniform(-6, -1))
        T = port.callcnvs_transitions(S, tp) if case % 2 == 0 else rng.dirichlet(np.ones(S), S)
        path, calls = port.c_hmm(T, ll, pos, 50000.0)
        want_path, want_calls = viterbi_py(T.tolist(), ll.tolist(), pos.tolist(), 50000.0)
        assert path.tolist() == want_path
        assert [tuple(int(v) for v in c) for c in calls] == want_calls


@pytest.mark.parametrize("S", [3, 5, 7])
def test_port_forward_agrees_with_50_digit_arithmetic(port, S):
    rng = np.random.default_rng(7 * S)
    for case in range(4):
        nobs = int(rng.integers(2, 60))
        ll, pos = random_case(rng, S, nobs, rough=case == 3)
        T = port.callcnvs_transitions(S, float(10 ** rng.uniform(-5, -2)))
        got = port.forward_loglik(T, ll, pos, 50000.0)
        want = forward_mp(T.tolist(), ll.tolist(), pos.tolist(), 50000.0)
        assert (got == want == -INF) or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (got, want)


@pytest.mark.parametrize("S", [5, 7])
def test_port_sstate_emission_agrees_with_the_mathematical_definition(port, S):
    """ll[state] = ln B(a1 + k, a2 + n - k) - ln B(a1, a2) with a1, a2 from (phi, e, odds[state]) as CNV_estimate.cpp:44-50,
    evaluated in 50-digit arithmetic.  The reference's own lnbeta carries ~1e-11 relative on these arguments."""
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(S)
    odds = port.state_odds(S)
    assert odds[1 if S == 3 else 2] == 1.0 and np.all(np.diff(odds) > 0) and odds[0] >= 0.05
    n = 40
    phi = 10 ** rng.uniform(-3.3, -2, n)
    e = rng.uniform(0.08, 0.3, n)
    tot = rng.integers(0, 4000, n).astype(np.int32)
    obs = rng.binomial(tot, e).astype(np.int32)
    for c in range(n):
        got = port.emission(phi[c], e[c], tot[c:c + 1], obs[c:c + 1], odds)[0]
        sd = mp.sqrt(mp.mpf(phi[c]) * mp.mpf(e[c]) * (1 - mp.mpf(e[c])))
        for s in range(S):
            es = mp.mpf(e[c]) * mp.mpf(odds[s]) / (mp.mpf(e[c]) * mp.mpf(odds[s]) + 1 - mp.mpf(e[c]))
            a1 = es * es * (1 - es) / (sd * sd) - es
            a2 = (1 - es) / es * a1
            want = float(mp.log(mp.beta(a1 + int(obs[c]), a2 + int(tot[c]) - int(obs[c]))) - mp.log(mp.beta(a1, a2)))
            assert abs(got[s] - want) <= 1e-9 * abs(want) + 1e-12, (c, s, got[s], want)
