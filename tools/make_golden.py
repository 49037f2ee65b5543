"""Generate the committed fixtures under tests/golden/ from the reference (run in the build container only).

  python tools/make_golden.py

Needs /root/reference (data/ExomeCount.RData, data/bedFiles/exons_hg19.bed) and the compiled reference
oracle/_ref/libexomedepth_ref.so (`make -C oracle ref`).  Outputs are small .npz/.json files that travel
with the repo; nothing at test time reads /root/reference.  Every OUTPUT array in ref_vectors.npz was
produced by the reference's own C/C++ (src/CNV_estimate.cpp, src/hmm.cpp, src/beta.c) compiled unmodified.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import rdata  # noqa: E402
from oracle import framing, ref  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def exomecount():
    g = rdata.load(f"{REF}/data/ExomeCount.RData")["ExomeCount"]
    rg = g.attr["ranges"]
    start = np.array(rg.attr["start"].value, np.int32)
    width = np.array(rg.attr["width"].value, np.int32)
    ld = g.attr["elementMetadata"].attr["listData"]
    cols = {nm: np.array(c.value) for nm, c in zip(ld.attr["names"].value, ld.value) if c.kind in ("int", "real")}
    out = dict(start=start, end=(start + width - 1).astype(np.int32))
    for k in ("Exome1", "Exome2", "Exome3", "Exome4"):
        assert np.all(cols[k] == np.round(cols[k]))
        out[k] = cols[k].astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "exomecount.npz"), **out)
    return out


def geometry():
    chrom, start, end = [], [], []
    with open(f"{REF}/data/bedFiles/exons_hg19.bed") as fh:
        for line in fh:
            c, s, e = line.split("\t")[:3]
            chrom.append(c)
            start.append(int(s))
            end.append(int(e))
    levels = list(dict.fromkeys(chrom))
    code = np.array([levels.index(c) for c in chrom], np.uint8)
    np.savez_compressed(os.path.join(OUT, "exons_hg19_geometry.npz"), chrom_code=code,
                        chrom_levels=np.array(levels), start=np.array(start, np.int32), end=np.array(end, np.int32))


def vectors(ec):
    rng = np.random.default_rng(20261017)
    ref.api().quiet(True)
    out = {}
    # --- lnbeta over every branch (beta.c:49-114): positive, near 1/2, small, negative, near negative integers
    x = np.concatenate([10 ** rng.uniform(-3, 5, 3000), rng.uniform(0.95, 2.05, 300), rng.uniform(-3, 0.5, 400),
                        -rng.integers(1, 40, 300) + rng.uniform(-0.014, 0.014, 300)])
    y = np.concatenate([10 ** rng.uniform(-3, 5, 3000), rng.uniform(0.95, 2.05, 300), rng.uniform(-3, 3, 400),
                        rng.uniform(0.3, 80, 300)])
    out["lnbeta_x"], out["lnbeta_y"], out["lnbeta_val"] = x, y, ref.lnbeta(x, y)
    # --- get_loglike_matrix: random per-bin phi/expected, mixtures 1 and 0.4, incl. pathological phi
    n = 6000
    phi = 10 ** rng.uniform(-3.5, -1.5, n)
    phi[-500:] = rng.uniform(0.3, 0.99, 500)
    e = rng.uniform(0.02, 0.6, n)
    tot = rng.poisson(10 ** rng.uniform(0, 3.7, n)).astype(np.int32)
    tot[::7] = 0
    obs = rng.binomial(tot, e).astype(np.int32)
    out.update(em_phi=phi, em_expected=e, em_total=tot, em_observed=obs,
               em_ll_mix1=ref.get_loglike_matrix(phi, e, tot, obs, 1.0),
               em_ll_mix04=ref.get_loglike_matrix(phi, e, tot, obs, 0.4))
    # --- C_hmm: random chains incl. negative distances (NaN transitions), -Inf emissions, ties
    hm = []
    for trial in range(40):
        nobs = int(rng.integers(2, 600))
        ll = -rng.exponential(5, (nobs, 3))
        if trial % 4 == 0:
            ll = np.round(ll)          # exact ties
        if trial % 3 == 0:
            ll[0] = [0, -np.inf, -np.inf]
            ll[-1] = [0, -100, -100]
        if trial % 5 == 0:
            ll[rng.integers(0, nobs)] = [0, -np.inf, -np.inf]
        pos = np.cumsum(rng.integers(-500, 30000, nobs)).astype(np.int32)
        tp = 10 ** rng.uniform(-6, -1)
        T = np.array([[1 - tp, tp / 2, tp / 2], [.5, .5, 0], [.5, 0, .5]]) if trial % 2 else rng.dirichlet(np.ones(3), 3)
        L = 50000.0 if trial % 4 else 1000.0
        path, calls = ref.c_hmm(T, ll, pos, L)
        hm.append((T, ll, pos, L, path, calls))
    out["hmm_n"] = np.array([len(hm)])
    for i, (T, ll, pos, L, path, calls) in enumerate(hm):
        out[f"hmm{i}_T"], out[f"hmm{i}_ll"], out[f"hmm{i}_pos"] = T, ll, pos
        out[f"hmm{i}_L"], out[f"hmm{i}_path"], out[f"hmm{i}_calls"] = np.array([L]), path.astype(np.int8), calls.astype(np.int32)
    # --- KAT-4: ExomeCount end to end
    test = ec["Exome4"].astype(float)
    reference = (ec["Exome1"] + ec["Exome2"] + ec["Exome3"]).astype(float)
    n = test.size
    phi, e = np.full(n, 0.0045127), np.full(n, 0.21763)
    ll = ref.get_loglike_matrix(phi, e, (test + reference).astype(np.int32), test.astype(np.int32), 1.0)
    res = framing.call_cnvs(ll, test, reference, e, ["chr1"] * n, ec["start"], ec["end"], ref.c_hmm)
    out["kat4_ll"] = ll
    out["kat4_path"] = res["paths"]["chr1"].astype(np.int8)
    keys = ("start_p", "end_p", "type", "nexons", "start", "end", "BF", "reads_expected", "reads_observed", "reads_ratio")
    out["kat4_calls"] = np.array([[c[k] for k in keys] for c in res["calls"]], float)
    out["kat4_cor"] = np.array([res["cor"]])
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **out)


def kats():
    """Known-answer vectors recorded in SURVEY.md §8c (KAT-1..3), checked here against the compiled reference."""
    kat2 = [(0.3, 5, 0.634215756099335), (0.01, 3, 4.5902323136238845), (1.2, 9, -2.7350379314936166),
            (1.005, 2.003, -0.70213639317424992), (5, 40, -15.507457058365487),
            (48.0084, 172.588, -116.46087412487566), (110.0084, 391.588, -265.16731573448669),
            (200, 1800, -651.84306537918803), (311, 2789, -1011.850479157375),
            (9000, 90000, -30162.559524536187), (30, 40, -48.30174909591608), (0.7, 0.9, 0.4398352519631657),
            (100000, 100003, -138636.00648789015), (20000, 300000, -74817.331317865435)]
    kat3 = [(400, 87, 1, -213.60390685644455, -210.00844068272443, -212.41617607605008),
            (400, 45, 1, -141.38680540872014, -146.99552374986257, -158.297460828672),
            (400, 130, 1, -265.17747519528342, -256.96056607220396, -253.05705633956973),
            (0, 0, 1, 0, 0, 0),
            (1000, 218, 1, -529.4370326827725, -525.22241078804689, -528.39220469007728),
            (37, 0, 1, -4.2534635963498317, -8.333509877386291, -11.97124827056814),
            (5000, 1088, 1, -2625.4378312133172, -2620.8514052876631, -2624.7600630650531),
            (400, 87, 0.5, -210.93515548129426, -210.00844068272443, -210.6336329792988),
            (400, 45, 0.5, -143.07355261309121, -146.99552374986257, -152.27298293642878),
            (37, 0, 0.5, -6.3533242590842036, -8.333509877386291, -10.202950470256269)]
    for x, y, v in kat2:
        got = float(ref.lnbeta(x, y))
        assert abs(got - v) <= 4e-15 * abs(v), (x, y, got, v)
    for tot, obs, mix, *vals in kat3:
        got = ref.get_loglike_matrix([0.0045127], [0.21763], [tot], [obs], mix)[0]
        assert np.allclose(got, vals, rtol=1e-14, atol=0), (tot, obs, got, vals)
    kat1 = dict(T=[[1 / 3] * 3] * 3, L=1.0, positions=list(range(1, 11)),
                loglik=[[0, -10, -10]] * 3 + [[-10, -10, 0]] * 3 + [[-10, 0, -10]] * 4,
                path=[0, 0, 0, 2, 2, 2, 1, 1, 1, 0], calls=[[4, 6, 2, 3], [4, 9, 1, 3]])
    p, c = ref.c_hmm(kat1["T"], kat1["loglik"], kat1["positions"], kat1["L"])
    assert p.tolist() == kat1["path"] and c.tolist() == kat1["calls"]
    json.dump(dict(source="SURVEY.md §8c; R/tools.R:74-86", kat1=kat1, kat2=kat2, kat3=kat3,
                   kat3_phi=0.0045127, kat3_expected=0.21763,
                   kat4=dict(colsums=[-6948400.4425085317, -6881204.9241145626, -6932337.0755078578],
                             ncalls=25, ndel=19, ndup=6, nbins_cnv=129)),
              open(os.path.join(OUT, "kat.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ec = exomecount()
    geometry()
    kats()
    vectors(ec)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
