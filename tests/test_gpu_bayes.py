"""GPU parity of the Gibbs sampler and the sample-based criteria against the reference."""
from itertools import islice

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Bm():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import bayes_pmf
    return bayes_pmf


def test_known_answer_table(Bm, golden):
    """SURVEY.md 8c Bayesian rows."""
    g = golden("known_answer_10x10_d2")
    R, U, V = g["ratings"], g["users"], g["items"]
    b = Bm.BayesianPMF(R, 2, subtract_mean=False)
    b.users, b.items = U.copy(), V.copy()
    sel = R[:, 0] == 0
    np.random.seed(0)
    x = b.sample_feature(0, True, np.zeros(2), np.eye(2), V, R[sel, 1].astype(int), R[sel, 2])
    np.testing.assert_allclose(x, [-28.04639559940905, 49.27729944075128], rtol=1e-10)
    np.random.seed(0)
    mu, alpha = b.sample_hyperparam(U, True)
    np.testing.assert_allclose(mu, [1.8946544328590074, 1.9165495158619739], rtol=1e-12)
    np.testing.assert_allclose(alpha, g["hyper_alpha"], rtol=1e-11)
    np.random.seed(0)
    s = list(islice(b.samples(num_gibbs=2), 3))
    np.testing.assert_allclose(s[2][0][0], [5.911244993127693, 17.59631201803792], rtol=1e-8)
    np.testing.assert_allclose(s[2][1][9], [-1.7869175224238054, 3.3705880307551848], rtol=1e-8)
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["samples_u"], rtol=1e-7, atol=1e-9)
    pv = b.pred_variance(s)
    np.testing.assert_allclose(pv, g["bayes_pred_variance"], rtol=1e-8)
    assert pv.sum() == pytest.approx(16816.729160266303, rel=1e-8)
    assert b.predict(s)[4, 7] == pytest.approx(84.793209488960429, rel=1e-9)
    np.testing.assert_array_equal(b.prob_ge_cutoff(s, .5), g["bayes_prob_ge_half"])


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-7), ("f32", 2e-3)])
def test_chain_subtract_mean(Bm, golden, dtype, tol):
    g = golden("gibbs_15x12_d3")
    b = Bm.BayesianPMF(g["ratings"], 3)
    b.compute_dtype = dtype
    b.users, b.items = g["users"].copy(), g["items"].copy()
    np.random.seed(int(g["seed"]))
    s = list(islice(b.samples(num_gibbs=2), 6))
    scale = np.abs(g["samples_u"]).max()
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["samples_u"], rtol=tol, atol=tol * scale)
    np.testing.assert_allclose(np.array([x[1] for x in s]), g["samples_v"], rtol=tol, atol=tol * scale)
    if dtype == "f64":
        ii, jj = g["cand_i"], g["cand_j"]
        which = (ii, jj)
        np.testing.assert_allclose(b.predict(s, which=which), g["bayes_predict"], rtol=1e-8)
        np.testing.assert_allclose(b.pred_variance(s, which=which), g["bayes_pred_variance"], rtol=1e-7)
        np.testing.assert_array_equal(b.prob_ge_cutoff(s, 3.5, which=which), g["bayes_prob_ge_3_5"])
        assert b.total_variance(s) == pytest.approx(float(g["bayes_total_variance"]), rel=1e-7)
        # selection as in bayes_pmf.full_test (:702-712)
        evals = b.pred_variance(s, which=which)
        assert int(np.argmax(evals)) == int(np.argmax(g["bayes_pred_variance"]))
        assert b.matrix_results(evals, which).shape == (15, 12)


def test_keys_registry(Bm):
    assert set(Bm.KEYS) == {'random', 'pred-variance', 'exp-variance', 'pred', 'prob-ge-3.5',
                            'prob-ge-.5', 'prob-ge-0'}
    k = Bm.KEYS['prob-ge-3.5']
    assert (k.key_fn, k.choose_max, k.wants_pool, k.args) == ('prob_ge_cutoff', True, False, (3.5,))


def test_active_loop_two_steps(Bm):
    """bayes_pmf.compare_active end to end on a tiny problem (pred-variance + exp-variance)."""
    import random
    np.random.seed(1); random.seed(1)
    n, m, d = 6, 5, 2
    u, v = np.random.normal(0, 1, (n, d)), np.random.normal(0, 1, (m, d))
    real = np.clip(np.round(u @ v.T + 3), 1, 5)
    known = [(i, (i * 2 + t) % m) for i in range(n) for t in range(2)] + [(0, 3), (1, 1)]
    known = sorted(set(known))
    ratings = np.array([(i, j, real[i, j]) for i, j in known], dtype=float)
    assert set(ratings[:, 1].astype(int)) == set(range(m))
    res = Bm.compare_active(['pred-variance'], d, real, ratings, rating_vals=(1, 2, 3, 4, 5),
                            num_steps=3, num_samps=8, threaded=False)
    steps = res['pred-variance']
    assert len(steps) == 3 and steps[1][0] == len(known) + 1 and steps[2][0] == len(known) + 2
    assert steps[1][3].shape == (n, m) and np.isfinite(steps[1][1])
    b = res['_initial_bpmf']
    samples = list(islice(b.samples(), 4))
    cand = sorted(b.unrated)[:2]
    which = tuple(np.array(cand).T)
    ev = b.exp_variance(samples, which=which, num_samps=4, fit_first=False)
    assert ev.shape == (2,) and np.all(np.isfinite(ev)) and np.all(ev > 0)
