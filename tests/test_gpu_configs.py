"""BASELINE configurations 2-4 at their real shapes, against fixtures made by running the
reference (its Cython build, its choose_training.py split script, its data files) in the build
container: tests/golden/make_golden_configs.py."""
import random
import time
from itertools import islice

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import active_pmf, bayes_pmf, scoring

    class NS:
        pass
    ns = NS()
    ns.A, ns.Bm, ns.S = active_pmf, bayes_pmf, scoring
    return ns


def _unknown_cells(R, n, m):
    known = np.zeros((n, m), bool)
    known[R[:, 0].astype(int), R[:, 1].astype(int)] = True
    return np.nonzero(~known)


def test_c2_drugbank_uv_entropy_over_all_unknowns(pkg, golden):
    """Config 2: drugbank 94 x 425 (+-1 interactions, 500 known), rank 5, the lookahead-entropy
    criterion over ALL 39,450 unknown cells -- infeasible in exact mode (k = 2595), one launch in
    scalable mode.  MAP fit trajectory, block posterior (KL as computed by the reference on the
    embedded matrix), all three lookahead criteria and the selected cell."""
    A = pkg.A
    g = golden("c2_drugbank")
    R = g["ratings"].astype(float)
    n, m, d = 94, 425, 5
    np.random.seed(0); random.seed(0)
    a = A.ActivePMF(R, d, rating_values=set(g["rating_vals"].tolist()), discrete_expectations=True)
    np.testing.assert_array_equal(a.users[:4], g["users0_head"])   # same constructor draws
    lls = list(a.fit_lls())
    assert len(lls) == int(g["fit_lls_steps"])
    assert lls[-1] == pytest.approx(float(g["fit_ll"]), rel=1e-10)
    np.testing.assert_allclose(a.users, g["users"], rtol=1e-7, atol=1e-9)
    assert a._use_blocks()                                      # k = 2595 > exact_max_dim
    a.users, a.items = g["users"].copy(), g["items"].copy()    # criteria from the reference's MAP
    a.blocks_tol, a.blocks_max_sweeps = 1e-12, 2000
    a.initialize_approx()
    a.fit_normal()
    assert a.kl_divergence() == pytest.approx(float(g["ref_kl_at_blocks"]), rel=1e-9)
    np.testing.assert_allclose(a.cov.A[:32], g["b_A_head"], rtol=1e-7, atol=1e-10)
    ii, jj = _unknown_cells(R, n, m)
    assert len(ii) == int(g["n_cand"]) == 39450
    pool = np.column_stack((ii, jj))
    t0 = time.perf_counter()
    ent = a._get_key_vals(pool, A.ActivePMF.exp_approx_entropy)
    dt = time.perf_counter() - t0
    np.testing.assert_allclose(ent, g["b_uv_entropy"], rtol=1e-9)
    want = int(np.argmin(g["b_uv_entropy"]))
    assert a.pick_query_point(pool, A.ActivePMF.exp_approx_entropy) == (int(ii[want]), int(jj[want]))
    sub = g["sub"]
    tv = a._get_key_vals(pool, A.ActivePMF.exp_total_variance)
    np.testing.assert_allclose(tv[sub], g["b_total_variance_sub"], rtol=1e-9)
    assert int(np.argmin(tv)) == int(g["b_total_variance_argmin"])
    np.testing.assert_allclose(a._get_key_vals(pool[sub], A.ActivePMF.exp_approx_entropy_byapprox),
                               g["b_uv_entropy_approx_sub"], rtol=1e-9)
    pv = a._get_key_vals(pool, A.ActivePMF.pred_variance)
    np.testing.assert_allclose(pv[sub], g["b_pred_var_sub"], rtol=1e-9)
    assert int(np.argmax(pv)) == int(g["b_pred_var_argmax"])
    np.testing.assert_allclose(pv[g["spots"]], g["ref_pred_var_spots"], rtol=1e-8)   # the reference's own
    print("c2: %d candidates x 2 values in %.1f ms (host call incl. copies); the numpy oracle took "
          "%.1f s" % (len(ii), dt * 1e3, float(g["oracle_lookahead_seconds"])))


def test_c3_movielens_shape(pkg, golden):
    """Config 3: 943 x 1682, 5,000 known, rank 10: objective, gradient and 40 line-search steps
    from the seeded start; `pred` on 2,048 unknown cells; block-posterior pred_variance and
    prob-ge-3.5 (checked with the reference on the 2-row models that hold the same blocks);
    selection over all 1,581,126 unknown cells through the class API."""
    A = pkg.A
    g = golden("c3_movielens")
    R = g["ratings"].astype(float)
    n, m, d = 943, 1682, 10
    np.random.seed(0); random.seed(0)
    a = A.ActivePMF(R, d, rating_values={1, 2, 3, 4, 5}, discrete_expectations=True, knowable=())
    np.testing.assert_array_equal(a.users[:4], g["users0_head"])
    assert a.log_likelihood() == pytest.approx(float(g["ll0"]), rel=1e-12)
    gu, gv = a.gradient()
    np.testing.assert_allclose(gu[:64], g["grad_u0_head"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gv[:64], g["grad_v0_head"], rtol=1e-10, atol=1e-12)
    assert gu.sum() == pytest.approx(float(g["grad_u0_sum"]), rel=1e-9)
    lls = list(islice(a.fit_lls(), 40))
    np.testing.assert_allclose(lls, g["lls40"], rtol=1e-11)
    np.testing.assert_allclose(a.users[:64], g["users_head"], rtol=1e-8, atol=1e-10)
    ii, jj = g["cand_i"], g["cand_j"]
    pool = np.column_stack((ii, jj))
    np.testing.assert_allclose(a._get_key_vals(pool, A.ActivePMF.pred), g["pred"], rtol=1e-8, atol=1e-10)
    assert a._use_blocks()
    a.blocks_tol, a.blocks_max_sweeps = 1e-9, 300
    a.initialize_approx()
    a.fit_normal()
    assert a.kl_divergence() == pytest.approx(float(g["b_kl"]), rel=1e-7)
    pv = a._get_key_vals(pool, A.ActivePMF.pred_variance)
    np.testing.assert_allclose(pv, g["b_pred_var"], rtol=1e-6)
    s = g["spots"]
    np.testing.assert_allclose(pv[s], g["ref_pred_var_spots"], rtol=1e-6)
    p35 = a._get_key_vals(pool, A.ActivePMF.prob_ge_3_5)
    np.testing.assert_allclose(p35[s], g["ref_prob_ge_3_5_spots"], rtol=1e-5, atol=1e-12)
    assert a.pick_query_point(pool, A.ActivePMF.pred_variance) == tuple(int(x) for x in pool[np.argmax(g["b_pred_var"])])
    # all unknown cells, as BASELINE words it: ndarray pool end to end, fused winner
    ai, aj = _unknown_cells(R, n, m)
    assert len(ai) == 1581126
    allpool = np.column_stack((ai, aj)).astype(np.int32)
    for key in (A.ActivePMF.pred, A.ActivePMF.pred_variance, A.ActivePMF.prob_ge_3_5):
        vals = a._get_key_vals(allpool, key)
        pick = a.pick_query_point(allpool, key)
        best = int(np.nanargmax(vals))
        assert pick == (int(ai[best]), int(aj[best]))


def test_c4_movielens_bayes(pkg, golden):
    """Config 4: BayesianPMF(rank 15, subtract_mean) on the same split: 30 line-search steps, 3
    Gibbs samples (num_gibbs=2) equal to the reference's chain, variance-based selection over all
    unrated cells equal to the reference's pick."""
    Bm = pkg.Bm
    g = golden("c4_movielens_bayes")
    R = g["ratings"].astype(float)
    n, m, d = 943, 1682, 15
    np.random.seed(0); random.seed(0)
    b = Bm.BayesianPMF(R, d, subtract_mean=True, rating_values={1, 2, 3, 4, 5}, knowable=())
    assert b.mean_rating == pytest.approx(float(g["mean_rating"]), rel=1e-14)
    lls = list(islice(b.fit_lls(), 30))
    np.testing.assert_allclose(lls, g["lls30"], rtol=1e-11)
    np.testing.assert_allclose(b.users[:64], g["users_head"], rtol=1e-8, atol=1e-10)
    np.random.seed(7)
    samples = list(islice(b.samples(num_gibbs=2), 3))
    for s, (us, vs) in enumerate(samples):
        np.testing.assert_allclose(us[:128], g["sample%d_u_head" % s], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(vs[:128], g["sample%d_v_head" % s], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose([us.sum(), vs.sum()], g["sample%d_sums" % s], rtol=1e-7)
    ii, jj = _unknown_cells(R, n, m)
    assert len(ii) == int(g["n_unrated"])
    var = b.pred_variance(samples, which=(ii, jj))
    pick = int(np.argmax(var))
    assert (int(ii[pick]), int(jj[pick])) == tuple(int(x) for x in g["pick"])
    assert var[pick] == pytest.approx(float(g["pick_value"]), rel=1e-6)
    s = g["spots"]
    np.testing.assert_allclose(var[s], g["var_spots"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(b.predict(samples, which=(ii[s], jj[s])), g["mean_spots"], rtol=1e-7)
    assert var.sum() == pytest.approx(float(g["var_sum"]), rel=1e-6)
