"""Pins the oracle's restatement of the sigma-learning / mini-batch fits, dense prediction, RMSE
and Wishart draws against tests/golden/extras_30x40_d4.npz,
produced by the reference's own Cython build (tests/golden/make_golden.py extras).  CPU only."""
import random

import numpy as np
import pytest

from oracle import pmf_oracle as O


@pytest.fixture(scope="module")
def g(golden):
    return golden("extras_30x40_d4")


def mean_of(g, sm):
    return dict(mean_rating=float(np.mean(g["ratings"][:, 2])) if sm else 0.0, subtract_mean=sm)


@pytest.mark.parametrize("tag", ["", "_sm", "_prior"])
def test_fit_with_sigmas_trajectory(g, tag):
    prior = {}
    if tag == "_prior":
        prior = dict(zip(("sig_u_mean", "sig_u_var", "sig_v_mean", "sig_v_var"), g["sig_prior"]))
    ref = g["ws_lls" + tag]
    U, V, lls, hyp = O.fit_with_sigmas_lls(g["ratings"], g["users0"], g["items0"], 5, 2,
                                           max_yields=len(ref), **mean_of(g, tag == "_sm"), **prior)
    # the trajectory is chaotic in the long run (accept/reject on differences of ~1e-2): the
    # first 150 accepted steps agree to rounding, the rest to the drift that rounding seeds
    np.testing.assert_allclose(lls[:150], ref[:150], rtol=1e-9)
    np.testing.assert_allclose(lls, ref, rtol=1e-5)
    np.testing.assert_allclose(hyp, g["ws_sigmas" + tag][-1], rtol=1e-4)
    np.testing.assert_allclose(U, g["ws_users" + tag], rtol=1e-3, atol=1e-5)


def test_old_objective_is_not_reevaluated(g):
    """the quirk of pmf_cy.pyx:285: re-evaluating old_ll under the new variances (the 'obvious'
    implementation) leaves the reference's trajectory within a few accepted steps"""
    ref = g["ws_lls"]
    R, U, V = g["ratings"], g["users0"], g["items0"]
    hyp = [1.0, 10.0, 10.0]
    lr, lls = 1e-4, []
    old = O.log_likelihood(R, U, V, *hyp)
    while len(lls) < 12:
        gu, gv = O.gradient(R, U, V, *hyp)
        while True:
            nu, nv = U + lr * gu, V + lr * gv
            new = O.log_likelihood(R, nu, nv, *hyp)
            if new > old:
                U, V, lr = nu, nv, lr * 1.25
                i = len(lls)
                if i % 5 == 0:
                    hyp[0] = O.update_sigma(R, U, V)
                if i % 2 == 0:
                    hyp[1], hyp[2] = O.update_sigma_uv(U, V, hyp[1], hyp[2])
                lls.append(new)
                old = O.log_likelihood(R, U, V, *hyp)          # <- what the reference does NOT do
                break
            lr *= .5
    assert not np.allclose(lls, ref[:12], rtol=1e-6)


@pytest.mark.parametrize("sm", [False, True])
def test_minibatch_validation(g, sm):
    tag = "_sm" if sm else ""
    R = g["ratings"].copy()
    np.random.seed(3); random.seed(3)
    # the split of pmf_cy.pyx:355-362
    total = R.shape[0]
    valid = set(random.sample(range(total), 40))
    train = R[tuple(i for i in range(total) if i not in valid), :]
    vidx = list(valid)
    mr = mean_of(g, sm)
    U, V, errs = O.fit_minibatches(R, g["users0"] * .3, g["items0"] * .3, 50, 6, lr=.05,
                                   train=train, **mr)
    np.testing.assert_allclose(errs, g["mb_errs" + tag][:, 0], rtol=1e-6)
    np.testing.assert_allclose(U, g["mb_users" + tag], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(V, g["mb_items" + tag], rtol=1e-9, atol=1e-12)
    pm = O.predicted_matrix(U, V, **mr)
    np.testing.assert_allclose(pm, g["pm" + tag], rtol=1e-9, atol=1e-12)
    vi, vj = R[vidx, 0].astype(int), R[vidx, 1].astype(int)
    assert O.rmse(pm[vi, vj], R[vidx, 2]) == pytest.approx(g["mb_errs" + tag][-1, 1], rel=1e-6)
    real, mask, rows = g["real"], g["mask"], g["rmse_rows"]
    got = [O.rmse(pm, real), O.rmse(pm, real, mask), O.rmse(pm, real, rows)]
    np.testing.assert_allclose(got, g["rmse3" + tag], rtol=1e-6)


def test_wishart_draws(g):
    np.random.seed(8)
    S = g["wishart_sigma"]
    np.testing.assert_allclose(O.sample_wishart(S, 7), g["wishart_direct"], rtol=1e-12)
    np.testing.assert_allclose(O.sample_wishart(S, 120), g["wishart_bartlett"], rtol=1e-12)
    # dof = 6.5 reaches the compiled reference as the C int 6 (bayes_pmf.pxd:7)
    np.testing.assert_allclose(O.sample_wishart(S, 6.5), g["wishart_bartlett_frac"], rtol=1e-12)


def test_bayes_rmse(g):
    samples = list(zip(g["br_samples_u"], g["br_samples_v"]))
    n, m = g["real"].shape
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    mr = float(np.mean(g["ratings"][:, 2]))
    pred = O.bayes_predict(samples, ii.ravel(), jj.ravel(), mr, True).reshape(n, m)
    got = [O.rmse(pred, g["real"]), O.rmse(pred, g["real"], g["mask"])]
    np.testing.assert_allclose(got, g["bayes_rmse"], rtol=1e-6)


def test_mirror_wishart_is_host_algebra(g):
    """the product's sample_wishart is d x d host algebra on the global numpy stream
    (bayes_pmf.py:41-59) -- no device needed; same draws as the reference"""
    from active_matrix_factorization_b200 import bayes_pmf
    np.random.seed(8)
    S = g["wishart_sigma"]
    np.testing.assert_allclose(bayes_pmf.sample_wishart(S, 7), g["wishart_direct"], rtol=1e-12)
    np.testing.assert_allclose(bayes_pmf.sample_wishart(S, 120), g["wishart_bartlett"], rtol=1e-12)
    np.testing.assert_allclose(bayes_pmf.sample_wishart(S, 6.5), g["wishart_bartlett_frac"], rtol=1e-12)


def test_pred_covs_entropy_bound_and_onestep(golden):
    """a14 / a17 rows: prediction covariance by Isserlis' theorem, its log-det bound, and the
    one-step lookahead utility, against the reference (tests/golden/more_criteria.npz)."""
    gl, c = golden("lookahead_6x7_d2"), golden("more_criteria")
    R, U, V, mean, cov = gl["ratings"], gl["users"], gl["items"], gl["mean"], gl["cov"]
    u, v = O.index_maps(6, 7, 2)
    np.testing.assert_allclose(O.pred_covs(u, v, mean, cov), c["pred_covs"], rtol=1e-9, atol=1e-11)
    assert O.pred_entropy_bound(u, v, mean, cov) == pytest.approx(float(c["pred_entropy_bound"]), rel=1e-10)
    unrated = list(zip(gl["cand_i"].tolist(), gl["cand_j"].tolist()))     # all unknown cells
    i, j = int(c["cand_i"][1]), int(c["cand_j"][1])
    one = O.lookahead_discrete(R, u, v, mean, cov, U, V, i, j, (0, 1), ("onestep", .5, unrated))
    assert one == pytest.approx(float(c["onestep_ge_half"][1]), rel=1e-6)
    one_a = O.lookahead_discrete(R, u, v, mean, cov, U, V, i, j, (0, 1), ("onestep", .5, unrated),
                                 use_map=False)
    assert one_a == pytest.approx(float(c["onestep_ge_half_approx"][1]), rel=1e-6)
    i, j = int(c["cand_i"][0]), int(c["cand_j"][0])
    peb = O.lookahead_discrete(R, u, v, mean, cov, U, V, i, j, (0, 1), "pred_entropy_bound")
    assert peb == pytest.approx(float(c["exp_pred_entropy_bound"][0]), rel=1e-6)


@pytest.mark.parametrize("tag", ["disc", "cont"])
def test_bayes_exp_variance(golden, tag):
    """a24: Bayesian lookahead from a seeded stream against the reference."""
    g, c = golden("gibbs_15x12_d3"), golden("more_criteria")
    samples = list(zip(g["samples_u"], g["samples_v"]))
    cells = list(zip(c["ev_cand_i"].tolist(), c["ev_cand_j"].tolist()))
    np.random.seed(5)
    ev = O.bayes_exp_variance(g["ratings"], g["users"], g["items"], samples, cells,
                              rating_values=(1, 2, 3, 4, 5) if tag == "disc" else None,
                              num_samps=3, num_integration_pts=5)
    np.testing.assert_allclose(ev, c["ev_" + tag], rtol=1e-6)
