"""Pins oracle/pmf_oracle.py (the numpy restatement) against fixtures produced by the
reference's own Cython build (tests/golden/make_golden.py) and against the SURVEY.md 8c
known-answer table.  CPU only."""
import numpy as np
import pytest

from oracle import pmf_oracle as O

RTOL = 1e-11


def hyp(g):
    return {k: float(g[k]) for k in ("sigma_sq", "sigma_u_sq", "sigma_v_sq") if k in g}


def test_known_answer_table(golden):
    """SURVEY.md section 8c values, quoted literally."""
    g = golden("known_answer_10x10_d2")
    R, U, V = g["ratings"], g["users"], g["items"]
    assert O.log_likelihood(R, U, V) == pytest.approx(-100332.69731950888, rel=1e-13)
    ll_full = O.log_likelihood(R, U, V) + O.ll_prior_adjustment(len(R), 10, 10, 2)
    assert ll_full == pytest.approx(-100378.74902136876, rel=1e-13)
    gu, gv = O.gradient(R, U, V)
    assert gu.sum() == pytest.approx(14100.289688036617, rel=1e-12)
    assert gv.sum() == pytest.approx(3006.2769559841336, rel=1e-12)
    np.testing.assert_allclose(gu[0], [4119.1451028768615, 4457.430367890687], rtol=1e-12)
    np.testing.assert_allclose(gv[3], [168.8811557711175, 209.6109163969808], rtol=1e-12)
    u, v = O.index_maps(10, 10, 2)
    mean, cov = g["mean"], g["cov"]
    assert O.kl_divergence(R, u, v, mean, cov) == pytest.approx(100087.39077408361, rel=1e-12)
    assert O.pred_mean_var(u, v, mean, cov, 1, 0)[1] == pytest.approx(115.09620486826407, rel=1e-11)
    e, var = O.pred_mean_var(u, v, mean, cov, 4, 7)
    assert e == pytest.approx(5.978942568676427, rel=1e-12)
    assert var == pytest.approx(20.169336740831291, rel=1e-11)
    gm, gc = O.normal_gradient(R, u, v, mean, cov)
    assert gm.sum() == pytest.approx(-16803.742482969799, rel=1e-11)
    assert gc.sum() == pytest.approx(-5721.9784928740974, rel=1e-11)
    assert gc[0, 1] == pytest.approx(179.88626295514206, rel=1e-11)
    assert np.linalg.slogdet(cov)[1] == pytest.approx(4.7926812458228989, rel=1e-12)


@pytest.mark.parametrize("name", ["known_answer_10x10_d2", "random_12x20_d5"])
def test_ll_grad_moments(golden, name):
    g = golden(name)
    R, U, V = g["ratings"], g["users"], g["items"]
    h = hyp(g)
    n, d = U.shape
    m = V.shape[0]
    assert O.log_likelihood(R, U, V, **h) == pytest.approx(float(g["ll"]), rel=RTOL)
    full = O.log_likelihood(R, U, V, **h) + O.ll_prior_adjustment(len(R), n, m, d, **h)
    assert full == pytest.approx(float(g["full_ll"]), rel=RTOL)
    gu, gv = O.gradient(R, U, V, **h)
    np.testing.assert_allclose(gu, g["grad_u"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(gv, g["grad_v"], rtol=RTOL, atol=1e-12)
    u, v = O.index_maps(n, m, d)
    mean, cov = g["mean"], g["cov"]
    ii, jj = g["cand_i"], g["cand_j"]
    mv = np.array([O.pred_mean_var(u, v, mean, cov, i, j) for i, j in zip(ii, jj)])
    np.testing.assert_allclose(mv[:, 0], g["pred_mean"], rtol=RTOL)
    np.testing.assert_allclose(mv[:, 1], g["pred_variance"], rtol=1e-9, atol=1e-10)
    closed = np.array([O.pred_mean_var_closed(*O._blocks(mean, cov, u, v, i, j))
                       for i, j in zip(ii, jj)])
    np.testing.assert_allclose(closed[:, 0], g["pred_mean"], rtol=RTOL)
    np.testing.assert_allclose(closed[:, 1], g["pred_variance"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(O.prob_ge_cutoff(mv[:, 0], mv[:, 1], .5), g["prob_ge_half"],
                               rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(O.prob_ge_cutoff(mv[:, 0], mv[:, 1], 3.5), g["prob_ge_3_5"],
                               rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(np.einsum("nd,nd->n", U[ii], V[jj]), g["pred"], rtol=RTOL)
    assert O.kl_divergence(R, u, v, mean, cov, **h) == pytest.approx(float(g["kl"]), rel=RTOL)
    gm, gc = O.normal_gradient(R, u, v, mean, cov, **h)
    np.testing.assert_allclose(gm, g["grad_mean"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(gc, g["grad_cov"], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("sm", [False, True])
def test_fit_trajectory(golden, sm):
    g = golden("fit_30x40_d4")
    tag = "_sm" if sm else ""
    R = g["ratings"]
    h = dict(mean_rating=float(g["mean_rating"]), subtract_mean=sm)
    assert O.log_likelihood(R, g["users0"], g["items0"], **h) == pytest.approx(float(g["ll0" + tag]), rel=RTOL)
    gu, gv = O.gradient(R, g["users0"], g["items0"], **h)
    np.testing.assert_allclose(gu, g["grad_u0" + tag], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(gv, g["grad_v0" + tag], rtol=RTOL, atol=1e-12)
    U, V, lls = O.fit_lls(R, g["users0"], g["items0"], **h)
    assert len(lls) == len(g["lls" + tag])
    np.testing.assert_allclose(lls, g["lls" + tag], rtol=1e-9)
    np.testing.assert_allclose(U, g["users_fit" + tag], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(V, g["items_fit" + tag], rtol=1e-7, atol=1e-9)
    assert O.update_sigma(R, U, V, **h) == pytest.approx(g["sigmas" + tag][0], rel=1e-8)


def test_variational_fit_and_lookahead(golden):
    g = golden("lookahead_6x7_d2")
    R, U, V = g["ratings"], g["users"], g["items"]
    u, v = O.index_maps(6, 7, 2)
    mean, cov, kls = O.fit_normal_kls(R, u, v, g["mean0"], g["cov0"])
    assert len(kls) == len(g["kls"])
    np.testing.assert_allclose(kls, g["kls"], rtol=1e-8)
    np.testing.assert_allclose(mean, g["mean"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(cov, g["cov"], rtol=1e-6, atol=1e-8)
    ii, jj = g["cand_i"], g["cand_j"]
    pv = [O.pred_mean_var(u, v, g["mean"], g["cov"], i, j)[1] for i, j in zip(ii, jj)]
    np.testing.assert_allclose(pv, g["pred_variance"], rtol=1e-9)
    for t in (0, 5, 11):       # a few candidates: each is a pair of full variational refits
        i, j = int(ii[t]), int(jj[t])
        ent = O.lookahead_discrete(R, u, v, g["mean"], g["cov"], U, V, i, j, (0, 1), "entropy")
        assert ent == pytest.approx(float(g["uv_entropy"][t]), rel=1e-6)
        ent2 = O.lookahead_discrete(R, u, v, g["mean"], g["cov"], U, V, i, j, (0, 1), "entropy",
                                    use_map=False)
        assert ent2 == pytest.approx(float(g["uv_entropy_approx"][t]), rel=1e-6)
        tv = O.lookahead_discrete(R, u, v, g["mean"], g["cov"], U, V, i, j, (0, 1), "total_variance")
        assert tv == pytest.approx(float(g["total_variance"][t]), rel=1e-6)


def test_gibbs_known_answer(golden):
    g = golden("known_answer_10x10_d2")
    R, U, V = g["ratings"], g["users"], g["items"]
    sel = R[:, 0] == 0
    np.random.seed(0)
    x = O.sample_feature(np.zeros(2), np.eye(2), V, R[sel, 1].astype(int), R[sel, 2])
    np.testing.assert_allclose(x, [-28.04639559940905, 49.27729944075128], rtol=1e-12)
    np.random.seed(0)
    mu, alpha = O.sample_hyperparam(U, np.eye(2), 2, 2, np.zeros(2))
    np.testing.assert_allclose(mu, [1.8946544328590074, 1.9165495158619739], rtol=1e-12)
    np.testing.assert_allclose(alpha, g["hyper_alpha"], rtol=1e-11)
    np.random.seed(0)
    s = O.gibbs_samples(R, U, V, 3, subtract_mean=False)
    np.testing.assert_allclose(s[2][0][0], [5.911244993127693, 17.59631201803792], rtol=1e-9)
    np.testing.assert_allclose(s[2][1][9], [-1.7869175224238054, 3.3705880307551848], rtol=1e-9)
    ii, jj = np.meshgrid(np.arange(10), np.arange(10), indexing="ij")
    pv = O.bayes_pred_variance(s, ii.ravel(), jj.ravel(), subtract_mean=False).reshape(10, 10)
    np.testing.assert_allclose(pv, g["bayes_pred_variance"], rtol=1e-8)
    assert pv.sum() == pytest.approx(16816.729160266303, rel=1e-9)
    pm = O.bayes_predict(s, ii.ravel(), jj.ravel(), subtract_mean=False).reshape(10, 10)
    assert pm[4, 7] == pytest.approx(84.793209488960429, rel=1e-9)
    pg = O.bayes_prob_ge_cutoff(s, ii.ravel(), jj.ravel(), .5, subtract_mean=False)
    np.testing.assert_array_equal(pg.reshape(10, 10), g["bayes_prob_ge_half"])


def test_gibbs_chain_subtract_mean(golden):
    g = golden("gibbs_15x12_d3")
    np.random.seed(int(g["seed"]))
    s = O.gibbs_samples(g["ratings"], g["users"], g["items"], 6,
                        mean_rating=float(g["mean_rating"]), subtract_mean=True)
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["samples_u"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.array([x[1] for x in s]), g["samples_v"], rtol=1e-8, atol=1e-10)
    ii, jj = g["cand_i"], g["cand_j"]
    kw = dict(mean_rating=float(g["mean_rating"]), subtract_mean=True)
    np.testing.assert_allclose(O.bayes_predict(s, ii, jj, **kw), g["bayes_predict"], rtol=1e-8)
    np.testing.assert_allclose(O.bayes_pred_variance(s, ii, jj, **kw), g["bayes_pred_variance"], rtol=1e-7)
    np.testing.assert_array_equal(O.bayes_prob_ge_cutoff(s, ii, jj, 3.5, **kw), g["bayes_prob_ge_3_5"])
