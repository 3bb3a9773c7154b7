"""GPU parity of the scalable-mode (block-diagonal) posterior and its lookahead
(csrc/blocks.cu) against oracle/block_oracle.py and the reference-made fixtures
(tests/golden/make_golden_configs.py), through the C ABI."""
import copy
import pickle

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from active_matrix_factorization_b200 import build
    build.build()
    import torch
    from active_matrix_factorization_b200 import _native as N, blocks, device as D, scoring
    from oracle import block_oracle as B

    class NS:
        pass
    ns = NS()
    ns.N, ns.blocks, ns.D, ns.S, ns.B, ns.torch = N, blocks, D, scoring, B, torch
    return ns


def _problem(seed, n, m, d, nnz, values=None):
    rng = np.random.RandomState(seed)
    cells = rng.permutation(n * m)[:nnz]
    ii, jj = cells // m, cells % m
    tu, tv = rng.normal(0, .6, (n, d)), rng.normal(0, .6, (m, d))
    r = np.einsum("nd,nd->n", tu[ii], tv[jj]) + rng.normal(0, .25, nnz)
    if values is not None:
        vals = np.array(sorted(values), float)
        r = vals[np.abs(r[:, None] - vals[None, :]).argmin(1)]
    R = np.column_stack((ii, jj, r)).astype(float)
    return R, rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))


def _fit_device(K, R, n, m, d, U, V, s2=1., su=10., sv=10., **kw):
    rat = K.D.Ratings.from_tuples(R, n, m, "f64")
    post = K.blocks.BlockPosterior(n, m, d, s2, su, sv)
    post.fit(rat, U, V, **kw)
    return rat, post


def test_one_call_fit_equals_the_python_driven_sweeps(K):
    """amf_blocks_fit (the whole coordinate descent in one library call) against fit_sweeps driven
    sweep by sweep from Python: same number of sweeps, identical tables"""
    n, m, d = 40, 55, 4
    R, U, V = _problem(7, n, m, d, 600)
    rat, _ = _fit_device(K, R, n, m, d, U, V, .7, 6., 11., sweeps=1, cov_term=False, update_mean=False)
    a = K.blocks.BlockPosterior(n, m, d, .7, 6., 11.)
    steps = len(list(a.fit_sweeps(rat, U, V, sweeps=200, tol=1e-9)))
    b = K.blocks.BlockPosterior(n, m, d, .7, 6., 11.)
    b.fit(rat, U, V, sweeps=200, tol=1e-9)
    assert b.sweeps_done == steps and 2 < steps < 200
    ha, hb = a.to_host(), b.to_host()
    for name in ("mean_u", "mean_v", "A", "B", "Lu", "Lv", "hu", "hv", "logdet_u", "logdet_v"):
        np.testing.assert_array_equal(getattr(ha, name), getattr(hb, name))
    c = K.blocks.BlockPosterior(n, m, d, .7, 6., 11.)
    c.fit(rat, U, V, sweeps=3, tol=0.0)              # the sweep limit stops it
    assert c.sweeps_done == 3


@pytest.mark.parametrize("n,m,d,nnz", [(6, 7, 2, 14), (12, 20, 5, 80), (30, 40, 10, 400),
                                        (20, 25, 16, 300), (10, 12, 32, 100), (9, 5, 1, 20)])
def test_fit_matches_oracle(K, n, m, d, nnz):
    """coordinate sweeps of the block-restricted KL: every table after 1 sweep (means fixed, no
    covariance term = initialize_approx) and after convergence; empty rows included"""
    R, U, V = _problem(d, n, m, d, nnz)
    R = R[R[:, 0] != 1]                          # user 1 has no rating: prior block only
    st1 = K.B.fit_blocks(R, n, m, d, U, V, .7, 6., 11., sweeps=1, cov_term=False, update_mean=False)
    rat, p1 = _fit_device(K, R, n, m, d, U, V, .7, 6., 11., sweeps=1, cov_term=False, update_mean=False)
    h = p1.to_host()
    np.testing.assert_allclose(h.A, st1.A, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(h.Lv, st1.Lv, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(h.mean_u, U, rtol=0, atol=0)
    st = K.B.fit_blocks(R, n, m, d, U, V, .7, 6., 11., sweeps=60, tol=1e-11)
    post = K.blocks.BlockPosterior(n, m, d, .7, 6., 11.)
    ri, rj, rr = (K.D.to_device(R[:, 0], np.int32), K.D.to_device(R[:, 1], np.int32),
                  K.D.to_device(R[:, 2], np.float64))
    kls = list(post.fit_sweeps(rat, U, V, sweeps=60, tol=1e-11, kl_of=lambda p: p.kl(ri, rj, rr)))
    h = post.to_host()
    for mine, want in ((h.mean_u, st.mu), (h.mean_v, st.mv), (h.A, st.A), (h.B, st.B),
                       (h.Lu, st.Lu), (h.Lv, st.Lv), (h.hu, st.hu), (h.hv, st.hv)):
        np.testing.assert_allclose(mine, want, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(h.logdet_u, np.linalg.slogdet(st.A)[1], rtol=1e-9, atol=1e-11)
    assert kls[-1] == pytest.approx(K.B.kl_blocks(st, R), rel=1e-9)
    assert np.all(np.diff(kls) <= 1e-9 * np.abs(kls[:-1]))
    assert post.entropy() == pytest.approx(K.B.entropy(st), rel=1e-9)
    assert post.total_variance() == pytest.approx(K.B.total_variance(st), rel=1e-9)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
@pytest.mark.parametrize("n,m,d", [(12, 20, 5), (30, 40, 10), (20, 25, 15), (10, 12, 32)])
def test_cell_criteria_match_oracle(K, n, m, d, dtype, tol):
    """approx mean / pred_variance (packed SDDMM for d(d+1) <= 128 / 256, block gather above) /
    prob_ge with the variance-as-scale quirk, fused arg-best"""
    R, U, V = _problem(100 + d, n, m, d, min(n * m // 2, 600))
    _rat, post = _fit_device(K, R, n, m, d, U, V, sweeps=8)
    h = post.to_host()
    st = K.B.Blocks(h.mean_u, h.mean_v, h.A, h.B, h.Lu, h.Lv, h.hu, h.hv, 1., 10., 10.)
    ai, aj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    ii, jj = ai.ravel(), aj.ravel()
    ci, cj = K.D.to_device(ii, np.int32), K.D.to_device(jj, np.int32)
    pm, pv = K.B.pred_mean_var(st, ii, jj)
    for crit, want in ((K.N.CRIT_APPROX_MEAN, pm), (K.N.CRIT_PRED_VARIANCE, pv)):
        got, best = post.score(crit, ci, cj, dtype)
        got = got.double().cpu().numpy()
        assert np.abs(got - want).max() <= tol * np.abs(want).max()
        bv, bi = K.S.unpack_best(best)
        assert bi == int(np.argmax(got)) and bv == pytest.approx(got.max(), rel=1e-12)
    from oracle import pmf_oracle as O
    got, best = post.score(K.N.CRIT_PROB_GE, ci, cj, dtype, cutoff=.5)
    want = O.prob_ge_cutoff(pm, pv, .5)
    assert np.abs(got.double().cpu().numpy() - want).max() <= max(tol, 1e-9) * 50


@pytest.mark.parametrize("name,n,m,d", [("blocks_6x7_d2", 6, 7, 2), ("blocks_12x20_d5", 12, 20, 5)])
def test_lookahead_matches_fixture(K, golden, name, n, m, d):
    """every mode of amf_blocks_lookahead against the oracle's fixture: raw evaluations, Delta-cdf
    weights from the MAP and from the approximation, 1 and 2 coordinate rounds, Gauss-Legendre
    window; fused arg-min; and the agreement with converged exact mode carried by the fixture"""
    g = golden(name)
    N = K.N
    R, U, V = g["ratings"], g["users"], g["items"]
    rat, post = _fit_device(K, R, n, m, d, U, V, sweeps=2000, tol=1e-13)
    assert post.kl(K.D.to_device(R[:, 0], np.int32), K.D.to_device(R[:, 1], np.int32),
                   K.D.to_device(R[:, 2], np.float64)) == pytest.approx(float(g["ref_kl_at_blocks"]), rel=1e-9)
    ii, jj = g["cand_i"], g["cand_j"]
    ci, cj = K.D.to_device(ii, np.int32), K.D.to_device(jj, np.int32)
    ev, _, _ = post.lookahead(N.LOOK_ENTROPY, ci, cj, [0., 1.], want_evals=True)
    np.testing.assert_allclose(ev.cpu().numpy(), g["b_entropy_evals"], rtol=1e-9)
    ev, _, _ = post.lookahead(N.LOOK_TOTAL_VARIANCE, ci, cj, [0., 1.], want_evals=True)
    np.testing.assert_allclose(ev.cpu().numpy(), g["b_tv_evals"], rtol=1e-9)
    bounds = np.array([-np.inf, .5, np.inf])
    mu_map = K.D.to_device(np.einsum("nk,nk->n", U[ii], V[jj]), np.float64)
    sd_map = K.torch.ones_like(mu_map)
    am, _ = post.score(N.CRIT_APPROX_MEAN, ci, cj, "f64")
    av, _ = post.score(N.CRIT_PRED_VARIANCE, ci, cj, "f64")
    for what, wname in ((N.LOOK_ENTROPY, "entropy"), (N.LOOK_TOTAL_VARIANCE, "total_variance")):
        for use_map in (True, False):
            for rounds in (1, 2):
                mu, sd = (mu_map, sd_map) if use_map else (am, av.sqrt())
                _, sc, best = post.lookahead(what, ci, cj, [0., 1.], N.WEIGHTS_DISCRETE, bounds, mu,
                                             sd, rounds=rounds)
                want = g["b_%s_%s_r%d" % (wname, "map" if use_map else "approx", rounds)]
                got = sc.cpu().numpy()
                np.testing.assert_allclose(got, want, rtol=1e-9)
                assert K.S.unpack_best(best)[1] == int(np.argmin(want))
    t, w = K.blocks.gauss_nodes(16)
    _, sc, _ = post.lookahead(N.LOOK_ENTROPY, ci, cj, t, N.WEIGHTS_NODES, w, mu_map, sd_map)
    np.testing.assert_allclose(sc.cpu().numpy(), g["b_entropy_nodes"], rtol=1e-9)
    # agreement with what exact mode converges to (thresholds as in tests/test_oracle_blocks.py)
    _, sc, _ = post.lookahead(N.LOOK_ENTROPY, ci, cj, [0., 1.], N.WEIGHTS_DISCRETE, bounds, mu_map, sd_map)
    sub = g["exact_sub"]
    rho = stats.spearmanr(sc.cpu().numpy()[sub], g["exact_entropy"])[0]
    assert rho >= (0.97 if d == 2 else 0.90)


def test_lookahead_large_d_and_ragged_pool(K):
    """d = 16 and 32 (16- and 32-lane groups), pool sizes that do not fill the last warp"""
    for n, m, d, ncand in ((20, 25, 16, 37), (10, 12, 32, 5), (30, 40, 10, 1)):
        R, U, V = _problem(7 + d, n, m, d, min(200, n * m - 10), values=(1, 2, 3, 4, 5))
        _rat, post = _fit_device(K, R, n, m, d, U, V, sweeps=6)
        h = post.to_host()
        st = K.B.Blocks(h.mean_u, h.mean_v, h.A, h.B, h.Lu, h.Lv, h.hu, h.hv, 1., 10., 10.)
        rng = np.random.RandomState(d)
        ii, jj = rng.randint(0, n, ncand), rng.randint(0, m, ncand)
        ci, cj = K.D.to_device(ii, np.int32), K.D.to_device(jj, np.int32)
        vals = [1., 2., 3., 4., 5.]
        for what, wname in ((K.N.LOOK_ENTROPY, "entropy"), (K.N.LOOK_TOTAL_VARIANCE, "total_variance")):
            for rounds in (1, 2):
                ev, _, _ = post.lookahead(what, ci, cj, vals, rounds=rounds, want_evals=True)
                want = K.B.lookahead_evals(st, ii, jj, vals, wname, rounds)
                np.testing.assert_allclose(ev.cpu().numpy(), want, rtol=1e-8)
    ci0 = K.torch.empty(0, dtype=K.torch.int32, device="cuda")
    _, sc, best = post.lookahead(K.N.LOOK_ENTROPY, ci0, ci0, [0.], K.N.WEIGHTS_DISCRETE,
                                 [-np.inf, np.inf], K.torch.empty(0, dtype=K.torch.float64, device="cuda"),
                                 K.torch.empty(0, dtype=K.torch.float64, device="cuda"))
    assert K.S.unpack_best(best)[1] == -1


def test_class_api_in_scalable_mode(K, golden):
    """ActivePMF with approx_mode='blocks': initialize_approx / fit_normal / kl_divergence /
    criteria / pick_query_point / deepcopy / pickle, values from the reference-made fixture"""
    from active_matrix_factorization_b200 import active_pmf as A
    g = golden("blocks_6x7_d2")
    a = A.ActivePMF(g["ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
    a.approx_mode = 'blocks'
    a.blocks_tol = 1e-13
    a.blocks_max_sweeps = 2000
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.initialize_approx()
    assert isinstance(a.cov, K.blocks.BlockDiagonal) and a.cov.shape == (26, 26)
    kls = list(a.fit_normal_kls())
    assert np.all(np.diff(kls) <= 1e-9 * np.abs(kls[:-1]))
    assert a.kl_divergence() == pytest.approx(float(g["ref_kl_at_blocks"]), rel=1e-9)
    assert a._approx_entropy() == pytest.approx(float(g["b_entropy0"]), rel=1e-9)
    assert a._total_variance() == pytest.approx(float(g["b_total_variance0"]), rel=1e-9)
    pool = list(zip(g["cand_i"].tolist(), g["cand_j"].tolist()))
    np.testing.assert_allclose(a._get_key_vals(pool, A.ActivePMF.pred_variance), g["b_pred_var"], rtol=1e-9)
    np.testing.assert_allclose(a._get_key_vals(pool, A.ActivePMF.exp_approx_entropy),
                               g["b_entropy_map_r1"], rtol=1e-9)
    np.testing.assert_allclose(a._get_key_vals(pool, A.ActivePMF.exp_total_variance_byapprox),
                               g["b_total_variance_approx_r1"], rtol=1e-9)
    assert a.pick_query_point(pool, A.ActivePMF.exp_approx_entropy) == pool[int(np.argmin(g["b_entropy_map_r1"]))]
    assert a.exp_approx_entropy(pool[3]) == pytest.approx(g["b_entropy_map_r1"][3], rel=1e-9)
    a.lookahead_rounds = 2
    np.testing.assert_allclose(a._get_key_vals(np.array(pool), A.ActivePMF.exp_total_variance),
                               g["b_total_variance_map_r2"], rtol=1e-9)
    a.lookahead_rounds = 1
    # the embedded matrix is the reference's layout: the exact-mode kernels agree on it
    b = A.ActivePMF(g["ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
    b.approx_mode = 'exact'
    b.users, b.items = a.users, a.items
    b.mean, b.cov = a.mean.copy(), a.cov.toarray()
    assert b.kl_divergence() == pytest.approx(a.kl_divergence(), rel=1e-10)
    np.testing.assert_allclose(b._get_key_vals(pool, A.ActivePMF.pred_variance), g["b_pred_var"], rtol=1e-8)
    for c in (copy.deepcopy(a), pickle.loads(pickle.dumps(a))):
        assert isinstance(c.cov, K.blocks.BlockDiagonal)
        assert c.kl_divergence() == pytest.approx(a.kl_divergence(), rel=1e-12)
    with pytest.raises(ValueError):
        a._get_key_vals(pool[:2], A.ActivePMF.exp_pred_entropy_bound)
    # continuous R_ij: fixed Gauss-Legendre window
    a.discrete_expectations = False
    np.testing.assert_allclose(a._get_key_vals(pool, A.ActivePMF.exp_approx_entropy),
                               g["b_entropy_nodes"], rtol=1e-9)
    # auto mode picks the family by dimension
    a.approx_mode = 'auto'
    assert not a._use_blocks()
    a.exact_max_dim = 10
    assert a._use_blocks()


def test_resident_candidate_pool(K, golden):
    """CandidatePool: device-resident pool across steps, O(1) removal, same winners as lists"""
    from active_matrix_factorization_b200 import active_pmf as A
    g = golden("blocks_12x20_d5")
    a = A.ActivePMF(g["ratings"], 5, rating_values={0, 1}, discrete_expectations=True)
    a.approx_mode = 'blocks'
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.initialize_approx()
    a.fit_normal()
    pool = sorted(a.unrated)
    cp = K.S.CandidatePool(np.array(pool))
    assert len(cp) == len(pool) and cp[0] == pool[0] and list(cp) == pool
    for key in (A.ActivePMF.pred, A.ActivePMF.pred_variance, A.ActivePMF.exp_approx_entropy):
        assert a.pick_query_point(cp, key) == a.pick_query_point(pool, key)
        np.testing.assert_allclose(a._get_key_vals(cp, key), a._get_key_vals(pool, key), rtol=1e-12)
    for _ in range(3):
        ij = a.pick_query_point(cp, A.ActivePMF.pred_variance)
        assert cp.remove(*ij) and not cp.remove(*ij)
        pool.remove(ij)
        assert len(cp) == len(pool) and set(cp) == set(pool)
        assert a.pick_query_point(cp, A.ActivePMF.pred_variance) == \
            max(pool, key=lambda c: a.pred_variance(c))
    ev = a.get_key_evals(cp, A.ActivePMF.pred)
    assert np.isfinite(ev).sum() == len(pool)
    # a pool the model has scored follows add_rating, like `unrated` does
    ij = a.pick_query_point(cp, A.ActivePMF.pred)
    a.add_rating(ij[0], ij[1], 1.0)
    assert ij not in set(cp) and len(cp) == len(pool) - 1 and ij not in a.unrated
