#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own Cython build (oracle/_ref,
built by oracle/build_ref.py from /root/reference) on seeded inputs.  Run in the build
container only (``/root/reference`` does not exist on the GPU box); the .npz outputs are
committed.  Everything is fp64.  Usage:  python tests/golden/make_golden.py
"""
import os
import pickle
import sys
from itertools import islice

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import build_ref, ref_loader  # noqa: E402

build_ref.build()
ref = ref_loader.load()
ActivePMF = ref.active_pmf.ActivePMF
BayesianPMF = ref.bayes_pmf.BayesianPMF
PMF = ref.pmf_cy.ProbabilisticMatrixFactorization
normal_gradient = ref.normal_exps_cy.normal_gradient


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def all_cells(n, m):
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    return ii.ravel(), jj.ravel()


def approx_outputs(a, cells):
    ii, jj = cells
    pv = np.array([a.pred_variance((i, j)) for i, j in zip(ii, jj)])
    pm = np.array([a.approx_pred_mean_var(i, j)[0] for i, j in zip(ii, jj)])
    ph = np.array([a.prob_ge_half((i, j)) for i, j in zip(ii, jj)])
    p35 = np.array([a.prob_ge_3_5((i, j)) for i, j in zip(ii, jj)])
    pr = np.array([a.pred((i, j)) for i, j in zip(ii, jj)])
    gm, gc = normal_gradient(a)
    return dict(cand_i=ii, cand_j=jj, pred=pr, pred_mean=pm, pred_variance=pv,
                prob_ge_half=ph, prob_ge_3_5=p35, kl=a.kl_divergence(),
                grad_mean=gm, grad_cov=gc, approx_entropy=a._approx_entropy())


# ---- case 1: the SURVEY.md 8c known-answer input (10x10, d=2, 18 ratings) ------------
def case_known_answer():
    with open("/root/reference/results/criteria/10x10_r1_u10_v10_1/data.pkl", "rb") as f:
        R = np.asarray(pickle.load(f)["_ratings"], dtype=float)
    n = m = 10
    d = 2
    users = (np.arange(n * d).reshape(n, d) + 1) / 7
    items = (np.arange(m * d).reshape(m, d)[::-1] + 2) / 3
    a = ActivePMF(R, d)
    a.users, a.items = users.copy(), items.copy()
    k = (n + m) * d
    mean = np.hstack((users.ravel(), items.ravel()))
    B = np.cos(np.arange(k * k).reshape(k, k) * 0.37)
    cov = B @ B.T / k + np.eye(k)
    a.mean, a.cov = mean.copy(), cov.copy()
    gu, gv = a.gradient()
    out = dict(ratings=R, users=users, items=items, mean=mean, cov=cov,
               ll=a.log_likelihood(), full_ll=a.full_ll(), grad_u=gu, grad_v=gv)
    out.update(approx_outputs(a, all_cells(n, m)))
    # Bayesian pieces on the same input
    b = BayesianPMF(R, 2, subtract_mean=False)
    b.users, b.items = users.copy(), items.copy()
    sel = R[:, 0] == 0
    np.random.seed(0)
    out["sample_feature0"] = b.sample_feature(0, True, np.zeros(2), np.eye(2), items,
                                              R[sel, 1].astype(int), R[sel, 2])
    np.random.seed(0)
    mu, alpha = b.sample_hyperparam(users, True)
    out["hyper_mu"], out["hyper_alpha"] = mu, alpha
    np.random.seed(0)
    s = list(islice(b.samples(num_gibbs=2), 3))
    out["samples_u"] = np.array([x[0] for x in s])
    out["samples_v"] = np.array([x[1] for x in s])
    out["bayes_pred_variance"] = b.pred_variance(s)
    out["bayes_predict"] = b.predict(s)
    out["bayes_prob_ge_half"] = b.prob_ge_cutoff(s, 0.5)
    save("known_answer_10x10_d2", **out)


def random_problem(seed, n, m, d, nnz, values=None, scale=1.0):
    rng = np.random.RandomState(seed)
    cells = rng.permutation(n * m)[:nnz]
    # every row and column rated at least once
    ii, jj = list(cells // m), list(cells % m)
    for i in range(n):
        if i not in ii:
            ii.append(i); jj.append(rng.randint(m))
    for j in range(m):
        if j not in jj:
            jj.append(j); ii.append(rng.randint(n))
    seen, keep = set(), []
    for t, ij in enumerate(zip(ii, jj)):
        if ij not in seen:
            seen.add(ij); keep.append(t)
    ii, jj = np.array(ii)[keep], np.array(jj)[keep]
    tu, tv = rng.normal(0, scale, (n, d)), rng.normal(0, scale, (m, d))
    r = np.einsum("nd,nd->n", tu[ii], tv[jj]) + rng.normal(0, .25, len(ii))
    if values is not None:
        vals = np.array(sorted(values), dtype=float)
        r = vals[np.abs(r[:, None] - vals[None, :]).argmin(1)]
    R = np.column_stack((ii, jj, r)).astype(float)
    users, items = rng.uniform(0, 1, (n, d)), rng.uniform(0, 1, (m, d))
    return rng, R, users, items


# ---- case 2: latent_d = 5 (exercises the d > 2 quirks of normal_gradient) ------------
def case_d5():
    n, m, d = 12, 20, 5
    rng, R, users, items = random_problem(11, n, m, d, 80)
    a = ActivePMF(R, d)
    a.users, a.items = users.copy(), items.copy()
    a.sigma_sq, a.sigma_u_sq, a.sigma_v_sq = 0.7, 6.0, 11.0
    k = (n + m) * d
    mean = np.hstack((users.ravel(), items.ravel())) + rng.normal(0, .05, k)
    S = rng.normal(0, 1, (k, k))
    cov = ref.active_pmf.project_psd(S @ S.T / k * .3 + rng.normal(0, .01, (k, k)), 1e-5)
    a.mean, a.cov = mean.copy(), cov.copy()
    gu, gv = a.gradient()
    out = dict(ratings=R, users=users, items=items, mean=mean, cov=cov,
               sigma_sq=a.sigma_sq, sigma_u_sq=a.sigma_u_sq, sigma_v_sq=a.sigma_v_sq,
               ll=a.log_likelihood(), full_ll=a.full_ll(), grad_u=gu, grad_v=gv)
    out.update(approx_outputs(a, all_cells(n, m)))
    save("random_12x20_d5", **out)


# ---- case 3: MAP fit trajectory + subtract_mean --------------------------------------
def case_fit():
    n, m, d = 30, 40, 4
    rng, R, users, items = random_problem(5, n, m, d, 300, scale=.8)
    real = rng.normal(0, 1, (n, m))
    out = dict(ratings=R, users0=users, items0=items, real=real)
    for sm in (False, True):
        p = PMF(R, d, sm)
        p.users, p.items = users.copy(), items.copy()
        tag = "_sm" if sm else ""
        gu, gv = p.gradient()
        out["ll0" + tag], out["grad_u0" + tag], out["grad_v0" + tag] = p.log_likelihood(), gu, gv
        lls = list(p.fit_lls())
        out["lls" + tag] = np.array(lls)
        out["users_fit" + tag], out["items_fit" + tag] = p.users, p.items
        out["mean_rating"] = p.mean_rating
        out["rmse" + tag] = p.rmse(real)
        p.update_sigma(); p.update_sigma_uv()
        out["sigmas" + tag] = np.array([p.sigma_sq, p.sigma_u_sq, p.sigma_v_sq])
    save("fit_30x40_d4", **out)


# ---- case 4: variational fit + lookahead criteria on a toy problem -------------------
def case_lookahead():
    n, m, d = 6, 7, 2
    rng, R, users, items = random_problem(3, n, m, d, 14, values=(0., 1.), scale=.7)
    a = ActivePMF(R, d, rating_values={0, 1}, discrete_expectations=True)
    a.users, a.items = users.copy(), items.copy()
    a.fit()
    np.random.seed(4)
    a.initialize_approx()
    cov0 = a.cov.copy()
    mean0 = a.mean.copy()
    kls = list(a.fit_normal_kls())
    out = dict(ratings=R, users=a.users, items=a.items, mean0=mean0, cov0=cov0,
               kls=np.array(kls), mean=a.mean, cov=a.cov)
    cand = sorted(a.unrated)
    ii, jj = np.array(cand).T
    out["cand_i"], out["cand_j"] = ii, jj
    out["pred_variance"] = np.array([a.pred_variance(c) for c in cand])
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out["uv_entropy"] = np.array([a.exp_approx_entropy(c) for c in cand])
        out["uv_entropy_approx"] = np.array([a.exp_approx_entropy_byapprox(c) for c in cand])
        out["total_variance"] = np.array([a.exp_total_variance(c) for c in cand])
    save("lookahead_6x7_d2", **out)


# ---- case 5: Gibbs chain with subtract_mean, d = 3 -----------------------------------
def case_gibbs():
    n, m, d = 15, 12, 3
    rng, R, users, items = random_problem(9, n, m, d, 70, values=(1., 2., 3., 4., 5.), scale=1.2)
    b = BayesianPMF(R, d)                       # subtract_mean=True default
    b.users, b.items = users.copy(), items.copy()
    np.random.seed(21)
    s = list(islice(b.samples(num_gibbs=2), 6))
    ii, jj = all_cells(n, m)
    out = dict(ratings=R, users=users, items=items, seed=21, mean_rating=b.mean_rating,
               samples_u=np.array([x[0] for x in s]), samples_v=np.array([x[1] for x in s]),
               cand_i=ii, cand_j=jj,
               bayes_predict=b.predict(s)[ii, jj],
               bayes_pred_variance=b.pred_variance(s)[ii, jj],
               bayes_prob_ge_3_5=b.prob_ge_cutoff(s, 3.5)[ii, jj],
               bayes_total_variance=b.total_variance(s))
    save("gibbs_15x12_d3", **out)


# ---- case 6: matrix-normal posterior -----------------------------------
def case_matrix_normal():
    MN = ref.mn_active_pmf.MNActivePMF
    mn_grad = ref.matrix_normal_exps_cy.matrixnormal_gradient
    import contextlib, io
    out = {}
    # (a) d = 3, random SPD Sigma / Omega: KL, gradient, criteria
    n, m, d = 9, 11, 3
    rng, R, users, items = random_problem(17, n, m, d, 45, values=(1., 2., 3., 4., 5.))
    a = MN(R, d)
    a.users, a.items = users.copy(), items.copy()
    a.sigma_sq, a.sigma_u_sq, a.sigma_v_sq = 0.8, 5.0, 12.0
    a.initialize_approx()
    s1, s2 = rng.normal(size=(n + m, n + m)), rng.normal(size=(d, d))
    a.mean = a.mean + rng.normal(0, .1, a.mean.shape)
    a.cov_useritems = s1 @ s1.T / (n + m) + np.eye(n + m) * .5
    a.cov_latents = s2 @ s2.T / d + np.eye(d) * .3
    gm, gs, go = mn_grad(a)
    ii, jj = all_cells(n, m)
    mv = np.array([a.approx_pred_mean_var(i, j) for i, j in zip(ii, jj)])
    out.update(a_ratings=R, a_users=users, a_items=items, a_mean=a.mean, a_sig=a.cov_useritems,
               a_om=a.cov_latents, a_hyp=np.array([a.sigma_sq, a.sigma_u_sq, a.sigma_v_sq]),
               a_kl=a.kl_divergence(), a_gm=gm, a_gs=gs, a_go=go, a_cand_i=ii, a_cand_j=jj,
               a_pred_mean=mv[:, 0], a_pred_var=mv[:, 1],
               a_prob_ge_3_5=np.array([a.prob_ge_3_5((i, j)) for i, j in zip(ii, jj)]),
               a_entropy=a._approx_entropy())
    # (b) toy fit + lookahead from the identity initialisation (deterministic)
    n, m, d = 5, 6, 2
    rng, R, users, items = random_problem(23, n, m, d, 12, values=(0., 1.), scale=.7)
    b = MN(R, d, rating_values={0, 1}, discrete_expectations=True)
    b.users, b.items = users.copy(), items.copy()
    b.fit()
    b.initialize_approx()
    kls = list(b.fit_normal_kls())
    cand = sorted(b.unrated)
    ci, cj = np.array(cand).T
    out.update(b_ratings=R, b_users=b.users, b_items=b.items, b_kls=np.array(kls), b_mean=b.mean,
               b_sig=b.cov_useritems, b_om=b.cov_latents, b_cand_i=ci, b_cand_j=cj,
               b_pred_var=np.array([b.pred_variance(c) for c in cand]),
               b_entropy=b._approx_entropy(), b_total_variance=b._total_variance())
    with contextlib.redirect_stdout(io.StringIO()):
        out["b_uv_entropy"] = np.array([b.exp_approx_entropy(c) for c in cand[:6]])
        out["b_exp_total_variance"] = np.array([b.exp_total_variance(c) for c in cand[:6]])
    save("matrix_normal", **out)


# ---- case 7: sigma-learning and mini-batch fits, dense prediction / RMSE, Bayes extras ---
def case_extras():
    import random as pyrandom
    n, m, d = 30, 40, 4
    rng, R, users, items = random_problem(5, n, m, d, 300, scale=.8)
    real = rng.normal(0, 1, (n, m))
    mask = rng.uniform(size=(n, m)) < .3
    out = dict(ratings=R, users0=users, items0=items, real=real, mask=mask)
    # (a) fit_with_sigmas_lls (pmf_cy.pyx:384-403), plain and with the log-normal variance prior
    for tag, sm, prior in (("", False, None), ("_sm", True, None), ("_prior", False, (0.5, 2.0, 1.0, 3.0))):
        p = PMF(R, d, sm)
        p.users, p.items = users.copy(), items.copy()
        if prior:
            p.sig_u_mean, p.sig_u_var, p.sig_v_mean, p.sig_v_var = prior
            out["sig_prior"] = np.array(prior)
        lls, sig = [], []
        for ll in islice(p.fit_with_sigmas_lls(5, 2), 400):
            lls.append(ll)
            sig.append((p.sigma_sq, p.sigma_u_sq, p.sigma_v_sq))
        out["ws_lls" + tag], out["ws_sigmas" + tag] = np.array(lls), np.array(sig)
        out["ws_users" + tag], out["ws_items" + tag] = p.users, p.items
    # (b) mini-batch SGD with a validation split (pmf_cy.pyx:308-381)
    for tag, sm in (("", False), ("_sm", True)):
        p = PMF(R.copy(), d, sm)
        p.users, p.items = users.copy() * .3, items.copy() * .3
        np.random.seed(3); pyrandom.seed(3)
        errs = list(islice(p.fit_minibatches_validation(50, 40, lr=.05), 6))
        out["mb_errs" + tag] = np.array(errs)
        out["mb_users" + tag], out["mb_items" + tag] = p.users, p.items
        q = PMF(R.copy(), d, sm)
        q.users, q.items = users.copy() * .3, items.copy() * .3
        np.random.seed(3); pyrandom.seed(3)
        q.fit_minibatches_until_validation(50, 40, lr=.05, stop_thresh=1e-3)
        out["mbu_users" + tag], out["mbu_items" + tag] = q.users, q.items
        # (c) dense prediction and RMSE on everything / a mask / an index tuple
        out["pm" + tag] = p.predicted_matrix()
        rows = np.array([0, 3, 29, 7])           # `on` is typed ndarray: a mask or row numbers
        out["rmse3" + tag] = np.array([p.rmse(real), p.rmse(real, mask), p.rmse(real, rows)])
        out["rmse_rows"] = rows
    # (d) Wishart draws, both schemes (bayes_pmf.py:41-59)
    S = rng.normal(size=(3, 3)); S = S @ S.T + np.eye(3)
    np.random.seed(8)
    out["wishart_sigma"] = S
    out["wishart_direct"] = ref.bayes_pmf.sample_wishart(S, 7)
    out["wishart_bartlett"] = ref.bayes_pmf.sample_wishart(S, 120)
    out["wishart_bartlett_frac"] = ref.bayes_pmf.sample_wishart(S, 6.5)
    # (e) bayes_rmse over a short chain (bayes_pmf.py:544-545)
    b = BayesianPMF(R, d)
    b.users, b.items = users.copy(), items.copy()
    np.random.seed(13)
    s = list(islice(b.samples(num_gibbs=2), 4))
    out["br_samples_u"] = np.array([x[0] for x in s])
    out["br_samples_v"] = np.array([x[1] for x in s])
    out["bayes_rmse"] = np.array([b.bayes_rmse(s, real), b.bayes_rmse(s, real, mask)])
    save("extras_30x40_d4", **out)


# ---- case 8: lookahead expectation by quadrature / Simpson (active_pmf.py:679-699) ------
def case_continuous():
    import contextlib, io
    g = np.load(os.path.join(HERE, "lookahead_6x7_d2.npz"))
    cand = list(zip(g["cand_i"].tolist(), g["cand_j"].tolist()))[:3]

    def model(values):
        a = ActivePMF(g["ratings"], 2, rating_values=values, discrete_expectations=False)
        a.users, a.items = g["users"].copy(), g["items"].copy()
        a.mean, a.cov = g["mean"].copy(), g["cov"].copy()
        return a

    # Adaptive quadrature (rating_values=None) is not pinned: the integrand is a line search
    # run to a stopping threshold, i.e. piecewise constant in its accepted-step count, so QUADPACK
    # subdivides to its limit (~1000 re-fits per candidate, ~20 min each in the reference) and
    # the value depends on where rounding puts the jumps.  Simpson over the rating values uses
    # the same re-fits with fixed nodes.
    with contextlib.redirect_stdout(io.StringIO()):
        a = model({0, .5, 1})
        simps = np.array([a._exp_with_rij(c, ActivePMF._approx_entropy, discretize='simps')
                          for c in cand])
        summed = np.array([a._exp_with_rij(c, ActivePMF._approx_entropy, discretize=True)
                           for c in cand])
    save("continuous_6x7_d2", cand_i=np.array([c[0] for c in cand]),
         cand_j=np.array([c[1] for c in cand]), simps_entropy=simps, summed_entropy=summed)


# ---- case 9: one-step lookahead, prediction-entropy bound, Bayesian expected variance ----
def case_more():
    import contextlib, io
    g = np.load(os.path.join(HERE, "lookahead_6x7_d2.npz"))
    a = ActivePMF(g["ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.mean, a.cov = g["mean"].copy(), g["cov"].copy()
    cand = list(zip(g["cand_i"].tolist(), g["cand_j"].tolist()))[:4]
    out = dict(cand_i=np.array([c[0] for c in cand]), cand_j=np.array([c[1] for c in cand]))
    out["pred_covs"] = a.approx_pred_covs()                       # active_pmf.py:324-390
    out["pred_entropy_bound"] = a._pred_entropy_bound()           # :559-574
    with contextlib.redirect_stdout(io.StringIO()):
        out["onestep_ge_half"] = np.array([a.onestep_ge_half(c) for c in cand])      # :459-500
        out["onestep_ge_half_approx"] = np.array([a.onestep_ge_half_approx(c) for c in cand])
        out["exp_pred_entropy_bound"] = np.array([a.exp_pred_entropy_bound(c) for c in cand[:2]])
    # Bayesian lookahead (bayes_pmf.py:457-525,560-602) from the chain of the gibbs fixture:
    # categorical fit over the rating values, and the normal fit integrated at 5 ppf points
    gg = np.load(os.path.join(HERE, "gibbs_15x12_d3.npz"))
    for tag, kw in (("disc", dict(rating_values=(1, 2, 3, 4, 5), discrete_expectations=True)),
                    ("cont", dict(rating_values=None, discrete_expectations=False,
                                  num_integration_pts=5))):
        b = BayesianPMF(gg["ratings"], 3, **kw)
        b.users, b.items = gg["users"].copy(), gg["items"].copy()
        samples = list(zip(gg["samples_u"], gg["samples_v"]))
        which = tuple(np.array(sorted(b.unrated)[:2]).T)
        np.random.seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            out["ev_" + tag] = b.exp_variance(samples, which=which, num_samps=3, fit_first=False)
        out["ev_cand_i"], out["ev_cand_j"] = which
    save("more_criteria", **out)


# ---- case 10: the reference's own test_normal_exps.py inputs, with its Cython answers ----
def case_moments():
    """Same distributions as check_expectation / test_exp_dotprod_sq of the reference's
    test_normal_exps.py (mean ~ N(0, 10), cov = project_psd(N(0, 5)^{dim x dim}, 1e-5)), three
    draws per function, evaluated by the compiled normal_exps_cy."""
    cy = ref.normal_exps_cy
    psd = ref.active_pmf.project_psd
    np.random.seed(42)
    out = {}
    for name, dim in (("tripexpect", 3), ("quadexpect", 4), ("exp_squared", 2), ("exp_a2bc", 3)):
        for t in range(3):
            mn = np.random.normal(0, 10, (dim,))
            cov = psd(np.random.normal(0, 5, (dim, dim)), 1e-5)
            out["%s_mean%d" % (name, t)], out["%s_cov%d" % (name, t)] = mn, cov
            out["%s_val%d" % (name, t)] = getattr(cy, name)(mn, cov, *range(dim))
    n, m, d = 1, 1, 3
    k = (n + m) * d
    u = np.arange(0, n * d).reshape(n, d).T
    v = np.arange(n * d, (n + m) * d).reshape(m, d).T
    for t in range(3):
        mn = np.random.normal(0, 10, (k,))
        cov = psd(np.random.normal(0, 5, (k, k)), 1e-5)
        out["exp_dotprod_sq_mean%d" % t], out["exp_dotprod_sq_cov%d" % t] = mn, cov
        out["exp_dotprod_sq_val%d" % t] = cy.exp_dotprod_sq(u, v, mn, cov, 0, 0)
    save("moments", **out)


if __name__ == "__main__":
    cases = dict(known_answer=case_known_answer, d5=case_d5, fit=case_fit,
                 lookahead=case_lookahead, gibbs=case_gibbs, matrix_normal=case_matrix_normal,
                 extras=case_extras, continuous=case_continuous, more=case_more, moments=case_moments)
    for name in (sys.argv[1:] or list(cases)):
        cases[name]()
