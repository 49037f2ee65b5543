"""Beta-binomial fit (stand-in for aod::betabin, R/class_definition.R:118-119) and the expected Bayes
factor of get.power.betabinom (R/tools.R:128-166).  Parity unpinned (aod / VGAM are third-party and absent): the GPU
fit is checked by the vanishing likelihood gradient (scipy digamma), against an independent scipy optimiser (1e-6
relative on mu and phi) and against the ExomeCount values SURVEY.md §8c records."""
import numpy as np
import pytest

from oracle import betabin as obb


def test_oracle_fit_recovers_the_survey_values(exomecount):
    ec = exomecount
    test, ref = ec["Exome4"], ec["Exome1"] + ec["Exome2"] + ec["Exome3"]
    mu, phi, ll = obb.fit(test, ref)
    assert abs(mu - 0.2176) < 2e-4 and abs(phi - 0.00451) < 2e-5          # SURVEY.md §8c "Data facts"
    a, b = mu * (1 - phi) / phi, (1 - mu) * (1 - phi) / phi
    gu, gw = obb.gradient_log(a, b, test, ref)
    assert abs(gu) < 1e-8 * abs(ll) and abs(gw) < 1e-8 * abs(ll) and ll == pytest.approx(obb.loglik(a, b, test, ref), rel=1e-12)


def test_expected_bf_mirror_matches_the_oracle():
    from exomedepth_b200 import betabin
    for size, phi, p, alt in ((200, 0.1, 0.2, 0.6), (200, 0.1, 0.2, 0.2), (731, 0.0045, 0.2176, 0.1221), (40, 0.02, 0.1, 0.05)):
        want = obb.get_power_betabinom(size, phi, p, alt)
        assert betabin.get_power_betabinom(size, phi, p, alt) == pytest.approx(want, rel=1e-10, abs=1e-13)
    assert obb.get_power_betabinom(200, 0.1, 0.2, 0.2) == pytest.approx(0.0, abs=1e-12)      # R/tools.R:123 example: no power
    assert obb.get_power_betabinom(200, 0.1, 0.2, 0.6) > 1.0


@pytest.mark.gpu
def test_cuda_fit_is_the_likelihood_maximiser(exomecount):
    from exomedepth_b200 import betabin, synth
    ec = exomecount
    test, ref = ec["Exome4"], ec["Exome1"] + ec["Exome2"] + ec["Exome3"]
    r = betabin.fit(test, ref)
    mu, phi = r["expected"][0], r["phi"][0]
    assert r["info"][0] >= 0 and abs(mu - 0.2176) < 2e-4 and abs(phi - 0.00451) < 2e-5
    a, b = mu * (1 - phi) / phi, (1 - mu) * (1 - phi) / phi
    gu, gw = obb.gradient_log(a, b, test, ref)
    scale = abs(obb.loglik(a, b, test, ref))
    assert abs(gu) < 1e-9 * scale and abs(gw) < 1e-9 * scale
    assert r["loglik"][0] == pytest.approx(obb.loglik(a, b, test, ref), rel=1e-11)
    # a cohort against one shared reference, and the same samples with a reference row each
    d = synth.cohort(9, n_bins=30000)
    got = betabin.fit(d["observed"], d["reference"])
    per = betabin.fit(d["observed"], np.tile(d["reference"], (9, 1)))
    assert np.array_equal(got["expected"], per["expected"]) and np.array_equal(got["phi"], per["phi"])
    for s in range(9):
        mu_o, phi_o, ll_o = obb.fit(d["observed"][s], d["reference"])
        assert got["info"][s] >= 0
        assert got["expected"][s] == pytest.approx(mu_o, rel=1e-6) and got["phi"][s] == pytest.approx(phi_o, rel=1e-5)
        assert got["loglik"][s] >= ll_o - 1e-9 * abs(ll_o)               # at least as good as the scipy optimum
        # the generating values: planted CNVs and the zero-inflated bins pull the fit a little, not far
        assert abs(got["expected"][s] / d["expected"][s] - 1) < 0.05


@pytest.mark.gpu
def test_cuda_fit_reports_degenerate_samples():
    from exomedepth_b200 import _lib, betabin
    ref = np.full(500, 100, np.int32)
    obs = np.zeros((3, 500), np.int32)
    obs[1] = 30
    obs[2, 7] = -3                                           # not a count
    r = betabin.fit(obs, ref)
    assert r["info"][0] == -1 and np.isnan(r["expected"][0])
    assert r["info"][2] == -2
    assert r["info"][1] in (-4, -3) or r["phi"][1] < 1e-6    # constant proportion: no over-dispersion to find


@pytest.mark.gpu
def test_cuda_expected_bf_batch_matches_the_oracle():
    """edb200_power_betabinom (one CTA per problem) against the scipy restatement of get.power.betabinom
    (R/tools.R:128-166, theory = FALSE, limit = FALSE), over the sizes and over-dispersions select.reference.set visits:
    median depths from a handful of reads to tens of thousands, phi from 1e-4 to 0.3."""
    from exomedepth_b200 import betabin
    rng = np.random.default_rng(5)
    size = np.concatenate([[0, 1, 2, 200, 200, 731, 40, 255, 256, 257], rng.integers(3, 40000, 40)]).astype(np.int32)
    phi = np.concatenate([[0.1] * 3, [0.1, 0.1, 0.0045, 0.02, 0.01, 0.01, 0.01], np.exp(rng.uniform(np.log(1e-4), np.log(0.3), 40))])
    p = np.concatenate([[0.2] * 3, [0.2, 0.2, 0.2176, 0.1, 0.3, 0.3, 0.3], rng.uniform(0.03, 0.6, 40)])
    odds = p / (1 - p) * 0.5
    alt = odds / (1 + odds)
    alt[3], alt[4] = 0.6, 0.2                                    # the two examples of R/tools.R:122-123
    got = betabin.get_power_betabinom_batch(size, phi, p, alt)
    for i in range(size.size):
        want = obb.get_power_betabinom(int(size[i]), float(phi[i]), float(p[i]), float(alt[i]))
        assert got[i] == pytest.approx(want, rel=2e-9, abs=1e-11), (i, size[i], phi[i], p[i], alt[i])
    assert abs(got[4]) < 1e-11 and got[3] > 1.0
    assert np.isnan(betabin.get_power_betabinom_batch([10], [0.0], [0.2], [0.1])[0])      # phi = 0: a / b are not finite
