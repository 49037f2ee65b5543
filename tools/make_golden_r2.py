"""Round-2 fixtures under tests/golden/ from the compiled reference (run in the build container only):

  python tools/make_golden_r2.py        ->  tests/golden/ref_vectors_r2.npz

Kept apart from tools/make_golden.py so that ref_vectors.npz stays byte-identical.  Every OUTPUT array was produced
by the reference's own C/C++ (oracle/_ref/libexomedepth_ref.so = src/CNV_estimate.cpp, src/hmm.cpp, src/beta.c and
the GSL chain under them, compiled unmodified) driven through the CallCNVs framing of oracle/framing.py.

  loo*      the four leave-one-out samples of data/ExomeCount.RData (test = Exome_s, reference = the other three,
            R/class_definition.R:354-409): likelihood matrix, Viterbi path, CNV.calls columns, cor(test, reference).
            phi / expected are per-sample scalars from the moment estimate below (NOT aod::betabin: third party,
            absent), passed to both sides, so they only have to be the same numbers.
  env*      get_loglike_matrix over an envelope of parameters: phi 1e-5 .. 0.99, expected 0.005 .. 0.97, totals up
            to 1e5 — where does 1e-10 relative hold against the reference?
  negint*   rows whose shape parameter a1 lands within 0.015 of -1 (expected -> 1, huge phi): the reference walks
            lngamma_sgn_sing (src/VP_gamma.c:795-894, psi/zeta chain) there.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import framing, ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KEYS = ("start_p", "end_p", "type", "nexons", "start", "end", "BF", "reads_expected", "reads_observed", "reads_ratio", "BF_raw",
        "reads_expected_raw")


def main():
    ref.api().quiet(True)
    rng = np.random.default_rng(20261018)
    ec = np.load(os.path.join(OUT, "exomecount.npz"))
    out = {}
    names = ["Exome1", "Exome2", "Exome3", "Exome4"]
    n = ec["start"].size
    for s, nm in enumerate(names):
        test = ec[nm].astype(float)
        reference = sum(ec[o].astype(float) for o in names if o != nm)
        # moment estimate of the beta-binomial: expected = pooled proportion, phi from the over-dispersion of the bins
        tot = test + reference
        e = float(test.sum() / tot.sum())
        keep = tot > 20
        z2 = (test[keep] - tot[keep] * e) ** 2 / (tot[keep] * e * (1 - e))
        phi = float(max((np.mean(z2) - 1) / (np.mean(tot[keep]) - 1), 1e-4))
        ll = ref.get_loglike_matrix(np.full(n, phi), np.full(n, e), tot.astype(np.int32), test.astype(np.int32), 1.0)
        res = framing.call_cnvs(ll, test, reference, np.full(n, e), ["chr1"] * n, ec["start"], ec["end"], ref.c_hmm)
        out[f"loo{s}_phi"], out[f"loo{s}_expected"] = np.array([phi]), np.array([e])
        out[f"loo{s}_ll"] = ll
        out[f"loo{s}_path"] = res["paths"]["chr1"].astype(np.int8)
        out[f"loo{s}_calls"] = np.array([[c[k] for k in KEYS] for c in res["calls"]], float)
        out[f"loo{s}_cor"] = np.array([res["cor"]])
        print(nm, "phi", phi, "expected", e, "calls", len(res["calls"]))
    out["loo_keys"] = np.array(KEYS)

    # ---- envelope
    m = 24000
    phi = 10 ** rng.uniform(-5, np.log10(0.99), m)
    e = np.concatenate([rng.uniform(0.005, 0.97, m // 2), rng.uniform(0.08, 0.35, m - m // 2)])
    tot = np.floor(10 ** rng.uniform(0, 5, m)).astype(np.int32)
    tot[::11] = 0
    p = np.clip(e * rng.choice([0.5, 1.0, 1.0, 1.0, 1.5], m), 0, 1)
    obs = rng.binomial(tot, p).astype(np.int32)
    obs[::13] = np.minimum(tot[::13], rng.integers(0, 3, obs[::13].size))
    out.update(env_phi=phi, env_expected=e, env_total=tot, env_observed=obs,
               env_ll=ref.get_loglike_matrix(phi, e, tot, obs, 1.0))

    # ---- shape parameters next to a negative integer
    k = 600
    e = rng.uniform(0.97, 0.9995, k)
    phi = 10 ** rng.uniform(2, 9, k)
    tot = rng.integers(0, 40, k).astype(np.int32)
    obs = np.minimum(tot, rng.integers(0, 40, k)).astype(np.int32)
    out.update(negint_phi=phi, negint_expected=e, negint_total=tot, negint_observed=obs,
               negint_ll=ref.get_loglike_matrix(phi, e, tot, obs, 1.0))
    sd2 = phi * e * (1 - e)
    a1 = e * e * (1 - e) / sd2 - e
    print("negint: a1 within 0.015 of -1 in", int(np.sum(np.abs(a1 + 1) < 0.015)), "of", k, "rows; NaN cells",
          int(np.isnan(out["negint_ll"]).sum()), "of", out["negint_ll"].size)
    np.savez_compressed(os.path.join(OUT, "ref_vectors_r2.npz"), **out)
    print("ref_vectors_r2.npz", os.path.getsize(os.path.join(OUT, "ref_vectors_r2.npz")))


if __name__ == "__main__":
    main()
