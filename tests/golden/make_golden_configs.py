#!/usr/bin/env python3
"""Round-2 fixtures: the scalable-mode (block) posterior pinned to the REFERENCE's own KL and to
the converged exact-mode optimum, and the BASELINE configurations C2-C4 at their real shapes.

Runs the reference's Cython build (oracle/_ref) and the reference's own data-preparation script
(choose_training.py) on the reference's own data files, in the build container only; the .npz
outputs are committed.  Usage:  python tests/golden/make_golden_configs.py [case ...]
"""
import contextlib
import io
import os
import random
import runpy
import sys
import tempfile
import time
from itertools import islice

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import block_oracle as B, build_ref, pmf_oracle as O, ref_loader  # noqa: E402

build_ref.build()
ref = ref_loader.load()
ActivePMF = ref.active_pmf.ActivePMF
BayesianPMF = ref.bayes_pmf.BayesianPMF
REF_ROOT = os.path.dirname(os.environ.get("AMF_REFERENCE_DIR", "/root/reference/python-pmf"))


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def choose_training(args, seed=0):
    """the reference's own split script, seeded"""
    random.seed(seed)
    np.random.seed(seed)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "split.npz")
        old = sys.argv
        sys.argv = ["choose_training.py"] + args + [out]
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                runpy.run_path(os.path.join(REF_ROOT, "choose_training.py"), run_name="__main__")
        except SystemExit:
            pass
        finally:
            sys.argv = old
        with np.load(out) as z:
            return {k: z[k] for k in z.files}


def block_state(st):
    return dict(b_mean_u=st.mu, b_mean_v=st.mv, b_A=st.A, b_B=st.B, b_Lu=st.Lu, b_Lv=st.Lv,
                b_hu=st.hu, b_hv=st.hv)


# ---- the block family against the reference's objective and its converged optimum -------------
def case_blocks(name, n, m, d, nnz, seed, n_exact_cands):
    sys.path.insert(0, HERE)
    from make_golden import random_problem
    rng, R, users, items = random_problem(seed, n, m, d, nnz, values=(0., 1.), scale=.7)
    a = ActivePMF(R, d, rating_values={0, 1}, discrete_expectations=True)
    a.users, a.items = users.copy(), items.copy()
    a.fit()
    U, V = a.users.copy(), a.items.copy()
    trace = []
    st = B.fit_blocks(R, n, m, d, U, V, sweeps=2000, tol=1e-13, trace=trace)
    mean, cov = st.embed()
    a.mean, a.cov = mean.copy(), cov.copy()
    out = dict(ratings=R, users=U, items=V, kl_trace=np.array(trace),
               ref_kl_at_blocks=a.kl_divergence(), oracle_kl_at_blocks=B.kl_blocks(st, R))
    out.update(block_state(st))
    if d <= 2:      # normal_gradient's l-sum is only right for d <= 2
        gm, gc = ref.normal_exps_cy.normal_gradient(a)
        mask = np.zeros_like(gc, bool)
        for b in range(n + m):
            mask[b * d:(b + 1) * d, b * d:(b + 1) * d] = True
        out["ref_grad_mean_max"] = np.abs(gm).max()
        out["ref_grad_cov_block_max"] = np.abs(gc[mask]).max()
    # the reference's own exact-mode fit from its random start, for the record
    np.random.seed(4)
    a.initialize_approx()
    kls = list(a.fit_normal_kls())
    out["ref_default_fit_kl"] = kls[-1]
    # converged optimum of the full-covariance objective, checked with the reference's KL
    fm, fc, fkl = B.fit_full_converged(R, n, m, d, mean, cov)
    a.mean, a.cov = fm.copy(), fc.copy()
    out["full_kl"], out["ref_kl_at_full"] = fkl, a.kl_divergence()
    fm2, fc2, fkl2 = B.fit_full_converged(R, n, m, d, a.mean, a.cov)
    cand = sorted(a.unrated)
    ii, jj = np.array(cand).T
    out["cand_i"], out["cand_j"] = ii, jj
    u, v = O.index_maps(n, m, d)
    out["exact_pred_var"] = np.array([O.pred_mean_var(u, v, fm, fc, i, j)[1] for i, j in cand])
    pm, pv = B.pred_mean_var(st, ii, jj)
    out["b_pred_mean"], out["b_pred_var"] = pm, pv
    # lookahead: block oracle over all candidates, converged exact mode over a subset
    for what in ("entropy", "total_variance"):
        for use_map in (True, False):
            for rounds in (1, 2):
                key = "b_%s_%s_r%d" % (what, "map" if use_map else "approx", rounds)
                out[key] = B.lookahead(st, ii, jj, what, use_map, {0, 1}, U, V, rounds=rounds)
    out["b_entropy_evals"] = B.lookahead_evals(st, ii, jj, [0., 1.], "entropy", 1)
    out["b_tv_evals"] = B.lookahead_evals(st, ii, jj, [0., 1.], "total_variance", 1)
    out["b_entropy_nodes"] = B.lookahead(st, ii, jj, "entropy", True, None, U, V, nq=16)
    out["b_entropy0"], out["b_total_variance0"] = B.entropy(st), B.total_variance(st)
    sub = np.arange(len(cand)) if n_exact_cands >= len(cand) else \
        np.sort(np.random.RandomState(1).permutation(len(cand))[:n_exact_cands])
    out["exact_sub"] = sub
    mu_map = np.einsum("nk,nk->n", U[ii[sub]], V[jj[sub]])
    vals, w = B.discrete_weights({0, 1}, mu_map, np.ones(len(sub)))
    ex_ent, ex_tv = np.zeros((len(sub), 2)), np.zeros((len(sub), 2))
    t0 = time.time()
    for c, t in enumerate(sub):
        for q, val in enumerate(vals):
            R2 = np.vstack((R, [ii[t], jj[t], val]))
            m2, c2, _ = B.fit_full_converged(R2, n, m, d, fm, fc)
            ex_ent[c, q] = np.linalg.slogdet(c2)[1]
            ex_tv[c, q] = O.pred_means_vars(u, v, m2, c2)[1].sum()
    print(name, "converged exact lookahead:", time.time() - t0, "s")
    out["exact_entropy"], out["exact_total_variance"] = (ex_ent * w).sum(1), (ex_tv * w).sum(1)
    save(name, **out)


# ---- C2: drugbank 94x425, 500 known, d=5, uv-entropy over all unknowns --------------------------
def case_c2():
    data = choose_training(["--drugbank", "--n-pick", "500", "--test-equal-classes", "--n-test",
                            "2000", os.path.join(REF_ROOT, "drugbank", "subset_94x425.npy")])
    R, real = data["_ratings"], data["_real"]
    n, m, d = 94, 425, 5
    np.random.seed(0)
    random.seed(0)
    a = ActivePMF(R, d, rating_values=set(data["_rating_vals"].tolist()), discrete_expectations=True)
    U0, V0 = a.users.copy(), a.items.copy()
    t0 = time.time()
    lls = list(a.fit_lls())
    fit_s = time.time() - t0
    U, V = a.users.copy(), a.items.copy()
    st = B.fit_blocks(R, n, m, d, U, V, sweeps=2000, tol=1e-12)
    out = dict(ratings=R.astype(np.int16), real=np.packbits(real > 0), users0_head=U0[:4], users=U,
               items=V, fit_lls_steps=len(lls), fit_ll=lls[-1], ref_fit_seconds=fit_s,
               rating_vals=data["_rating_vals"], b_A_head=st.A[:32], b_mean_v_head=st.mv[:32])
    # the reference's own KL at the block posterior, on the embedded 2595 x 2595 covariance
    mean, cov = st.embed()
    a.mean, a.cov = mean, cov
    t0 = time.time()
    out["ref_kl_at_blocks"] = a.kl_divergence()
    out["ref_kl_seconds"] = time.time() - t0
    out["oracle_kl_at_blocks"] = B.kl_blocks(st, R)
    cand = sorted(a.unrated)
    ii, jj = np.array(cand).T
    out["n_cand"] = len(cand)
    t0 = time.time()
    out["b_uv_entropy"] = B.lookahead(st, ii, jj, "entropy", True, {-1, 1}, U, V)
    out["oracle_lookahead_seconds"] = time.time() - t0
    sub = np.sort(np.random.RandomState(3).permutation(len(cand))[:4096])
    out["sub"] = sub
    tv = B.lookahead(st, ii, jj, "total_variance", True, {-1, 1}, U, V)
    out["b_total_variance_sub"], out["b_total_variance_argmin"] = tv[sub], np.argmin(tv)
    out["b_uv_entropy_approx_sub"] = B.lookahead(st, ii[sub], jj[sub], "entropy", False, {-1, 1}, U, V)
    pm, pv = B.pred_mean_var(st, ii, jj)
    out["b_pred_var_sub"], out["b_pred_var_argmax"] = pv[sub], np.argmax(pv)
    # spot check of the per-candidate criterion with the REFERENCE: pred_variance touches only
    # the 2d x 2d sub-block of cov on (u_i, v_j) (normal_exps_cy.pyx:127-134)
    spots = np.random.RandomState(2).permutation(len(cand))[:64]
    out["spots"] = spots
    out["ref_pred_var_spots"] = np.array([a.pred_variance(cand[t]) for t in spots])
    out["ref_pred_mean_spots"] = np.array([a.approx_pred_mean_var(*cand[t])[0] for t in spots])
    save("c2_drugbank", **out)


def movielens_split():
    return choose_training(["--pick-known-frac", "0.05", "--test-at-random", "--test-known-frac",
                            "0.05", os.path.join(REF_ROOT, "movielens-100k", "ratings_matrix.npy.gz")])


# ---- C3: movielens-100k shape, 5000 known, d=10 -------------------------------------------------
def case_c3():
    data = movielens_split()
    R = data["_ratings"]
    n, m, d = 943, 1682, 10
    np.random.seed(0)
    random.seed(0)
    a = ActivePMF(R, d, rating_values=set(data["_rating_vals"].tolist()), discrete_expectations=True,
                  knowable=())
    # the constructor's draws (pmf_cy.pyx:74-75) are the first of the seeded stream: tests rebuild
    # the same start with np.random.seed(0) instead of storing it
    out = dict(ratings=R.astype(np.int16), ll0=a.log_likelihood(), users0_head=a.users[:4].copy())
    gu, gv = a.gradient()
    out["grad_u0_head"], out["grad_v0_head"] = gu[:64], gv[:64]
    out["grad_u0_sum"], out["grad_v0_sum"] = gu.sum(), gv.sum()
    t0 = time.time()
    lls = list(islice(a.fit_lls(), 40))
    out["ref_fit40_seconds"] = time.time() - t0
    out["lls40"] = np.array(lls)
    U, V = a.users.copy(), a.items.copy()
    out["users_head"], out["items_head"] = U[:64], V[:64]
    rng = np.random.RandomState(5)
    rated = set(zip(R[:, 0].astype(int), R[:, 1].astype(int)))
    cells = rng.permutation(n * m)[:4000]
    cand = [(c // m, c % m) for c in cells if (c // m, c % m) not in rated][:2048]
    ii, jj = np.array(cand).T
    out["cand_i"], out["cand_j"] = ii, jj
    out["pred"] = np.array([a.pred(c) for c in cand])
    # block posterior (scalable mode); per-candidate criteria checked with the reference on the
    # 2-row model that holds the same blocks (k = 2d)
    st = B.fit_blocks(R, n, m, d, U, V, sweeps=300, tol=1e-9)
    out["b_kl"] = B.kl_blocks(st, R)
    pm, pv = B.pred_mean_var(st, ii, jj)
    out["b_pred_mean"], out["b_pred_var"] = pm, pv
    tiny = ActivePMF(np.array([[0., 0., 3.]]), d)
    spots = np.arange(96)
    ref_pv, ref_pm, ref_p35 = [], [], []
    for t in spots:
        i, j = cand[t]
        tiny.mean = np.hstack((st.mu[i], st.mv[j]))
        cov = np.zeros((2 * d, 2 * d))
        cov[:d, :d], cov[d:, d:] = st.A[i], st.B[j]
        tiny.cov = cov
        ref_pv.append(tiny.pred_variance((0, 0)))
        ref_pm.append(tiny.approx_pred_mean_var(0, 0)[0])
        ref_p35.append(tiny.prob_ge_3_5((0, 0)))
    out["spots"] = spots
    out["ref_pred_var_spots"], out["ref_pred_mean_spots"] = np.array(ref_pv), np.array(ref_pm)
    out["ref_prob_ge_3_5_spots"] = np.array(ref_p35)
    out["b_prob_ge_3_5"] = O.prob_ge_cutoff(pm, pv, 3.5)
    save("c3_movielens", **out)


# ---- C4: same split, BayesianPMF d=15, 3 Gibbs samples, variance pick over all unrated ----------
def case_c4():
    data = movielens_split()
    R = data["_ratings"]
    n, m, d = 943, 1682, 15
    np.random.seed(0)
    random.seed(0)
    b = BayesianPMF(R, d, subtract_mean=True, rating_values=set(data["_rating_vals"].tolist()),
                    knowable=())
    t0 = time.time()
    lls = list(islice(b.fit_lls(), 30))
    out = dict(ratings=R.astype(np.int16), lls30=np.array(lls), users_head=b.users[:64].copy(),
               items_head=b.items[:64].copy(), mean_rating=b.mean_rating,
               ref_fit30_seconds=time.time() - t0)
    np.random.seed(7)
    t0 = time.time()
    samples = list(islice(b.samples(num_gibbs=2), 3))
    out["ref_gibbs3_seconds"] = time.time() - t0
    for s, (us, vs) in enumerate(samples):      # a chain that matches on 128 rows matches
        out["sample%d_u_head" % s], out["sample%d_v_head" % s] = us[:128], vs[:128]
        out["sample%d_sums" % s] = np.array([us.sum(), vs.sum()])
    known = np.zeros((n, m), bool)
    known[R[:, 0].astype(int), R[:, 1].astype(int)] = True
    ii, jj = np.nonzero(~known)
    t0 = time.time()
    var = b.pred_variance(samples, which=(ii, jj))
    out["ref_variance_seconds"] = time.time() - t0
    out["n_unrated"] = len(ii)
    out["pick"] = np.array([ii[np.argmax(var)], jj[np.argmax(var)]])
    out["pick_value"] = var.max()
    rng = np.random.RandomState(9)
    spots = np.sort(rng.permutation(len(ii))[:2048])
    out["spots"] = spots
    out["var_spots"] = var[spots]
    out["mean_spots"] = b.predict(samples, which=(ii[spots], jj[spots]))
    out["var_sum"] = var.sum()
    save("c4_movielens_bayes", **out)


# ---- adaptive quadrature of the continuous lookahead, pinned at the nodes SciPy visits first ------
def case_quad_nodes():
    """active_pmf.py:691-699: stats.norm.expect(calculate_fn, lb, ub, epsrel=.02) = QUADPACK's
    adaptive 21-point Gauss-Kronrod.  The integrand is a line search run to a stopping threshold
    (piecewise constant in its accepted-step count), so the adaptive loop subdivides for ~1000
    evaluations; what CAN be pinned is the integrand at the nodes of the first pass: the first
    21 values of v the reference asks for, and what it computes there."""
    g = np.load(os.path.join(HERE, "lookahead_6x7_d2.npz"))
    cand = list(zip(g["cand_i"].tolist(), g["cand_j"].tolist()))[:2]
    a = ActivePMF(g["ratings"], 2, rating_values=None, discrete_expectations=False)
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.mean, a.cov = g["mean"].copy(), g["cov"].copy()

    class Enough(Exception):
        pass
    out = {"cand_i": np.array([c[0] for c in cand]), "cand_j": np.array([c[1] for c in cand])}
    steps = []

    def counting_fit(self):          # fit_normal (active_pmf.py:242-249) with its step count kept
        steps.append(sum(1 for _kl in self.fit_normal_kls()))
    ActivePMF.fit_normal = counting_fit
    for t, c in enumerate(cand):
        seen = []
        del steps[:]

        def fn(model, v=None, seen=seen):
            val = ActivePMF._approx_entropy(model)
            seen.append((v, val))
            if len(seen) >= 21:
                raise Enough()
            return val
        fn.__name__ = "_approx_entropy"
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                a._exp_with_rij(c, fn, use_map=True, pass_v=True)
        except Enough:
            pass
        out["nodes%d" % t] = np.array([s[0] for s in seen])
        out["values%d" % t] = np.array([s[1] for s in seen])
        out["steps%d" % t] = np.array(steps[:len(seen)])
    save("quad_nodes_6x7_d2", **out)


CASES = {
    "quad_nodes": case_quad_nodes,
    "blocks_6x7": lambda: case_blocks("blocks_6x7_d2", 6, 7, 2, 14, 3, 1000),
    "blocks_12x20": lambda: case_blocks("blocks_12x20_d5", 12, 20, 5, 80, 11, 40),
    "c2": case_c2, "c3": case_c3, "c4": case_c4,
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        CASES[name]()
        print(name, "done in %.1f s" % (time.time() - t0))
