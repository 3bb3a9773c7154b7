"""GPU parity of the matrix-normal variant against the reference's golden
outputs (tests/golden/matrix_normal.npz)."""
import copy
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import mn_active_pmf
    return mn_active_pmf


def test_mn_kl_gradient_criteria(M, golden):
    g = golden("matrix_normal")
    a = M.MNActivePMF(g["a_ratings"], 3)
    a.users, a.items = g["a_users"].copy(), g["a_items"].copy()
    a.sigma_sq, a.sigma_u_sq, a.sigma_v_sq = g["a_hyp"]
    a.mean, a.cov_useritems, a.cov_latents = g["a_mean"].copy(), g["a_sig"].copy(), g["a_om"].copy()
    assert a.kl_divergence() == pytest.approx(float(g["a_kl"]), rel=1e-11)
    gm, gs, go = M.matrixnormal_gradient(a)
    for got, ref in ((gm, g["a_gm"]), (gs, g["a_gs"]), (go, g["a_go"])):
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    pool = list(zip(g["a_cand_i"].tolist(), g["a_cand_j"].tolist()))
    pv = np.array(a._get_key_vals(pool, M.MNActivePMF.pred_variance, None, None))
    np.testing.assert_allclose(pv, g["a_pred_var"], rtol=1e-9)
    pg = np.array(a._get_key_vals(pool, M.MNActivePMF.prob_ge_3_5, None, None))
    np.testing.assert_allclose(pg, g["a_prob_ge_3_5"], rtol=1e-8, atol=1e-300)
    mn, var = a.approx_pred_mean_var(2, 5)
    t = pool.index((2, 5))
    assert mn == pytest.approx(float(g["a_pred_mean"][t]), rel=1e-12)
    assert var == pytest.approx(float(g["a_pred_var"][t]), rel=1e-9)
    assert a._approx_entropy() == pytest.approx(float(g["a_entropy"]), rel=1e-11)
    assert a.pick_query_point(pool, M.MNActivePMF.pred_variance) == pool[int(np.argmax(g["a_pred_var"]))]
    assert M.exp_dotprod_sq(9, a.mean, a.cov_useritems, a.cov_latents, 2, 5) == pytest.approx(
        float(g["a_pred_var"][t] + g["a_pred_mean"][t] ** 2), rel=1e-10)
    assert len(M.KEY_FUNCS) == 13
    with pytest.raises(NotImplementedError):
        a.approx_pred_covs()


def test_mn_fit_and_lookahead(M, golden):
    g = golden("matrix_normal")
    b = M.MNActivePMF(g["b_ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
    b.users, b.items = g["b_users"].copy(), g["b_items"].copy()
    b.initialize_approx()
    kls = list(b.fit_normal_kls())
    assert len(kls) == len(g["b_kls"])
    np.testing.assert_allclose(kls, g["b_kls"], rtol=1e-8)
    np.testing.assert_allclose(b.cov_useritems, g["b_sig"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(b.cov_latents, g["b_om"], rtol=1e-6, atol=1e-8)
    assert b._approx_entropy() == pytest.approx(float(g["b_entropy"]), rel=1e-7)
    assert b._total_variance() == pytest.approx(float(g["b_total_variance"]), rel=1e-7)
    pool = list(zip(g["b_cand_i"].tolist(), g["b_cand_j"].tolist()))
    pv = np.array(b._get_key_vals(pool, M.MNActivePMF.pred_variance, None, None))
    np.testing.assert_allclose(pv, g["b_pred_var"], rtol=1e-6)
    # lookahead criteria: one launch of 2 * |pool| matrix-normal re-fits each
    b.mean, b.cov_useritems, b.cov_latents = g["b_mean"].copy(), g["b_sig"].copy(), g["b_om"].copy()
    ent = np.array(b._get_key_vals(pool[:6], M.MNActivePMF.exp_approx_entropy, None, None))
    np.testing.assert_allclose(ent, g["b_uv_entropy"], rtol=1e-5)
    tv = np.array(b._get_key_vals(pool[:6], M.MNActivePMF.exp_total_variance, None, None))
    np.testing.assert_allclose(tv, g["b_exp_total_variance"], rtol=1e-5)
    for c in (copy.deepcopy(b), pickle.loads(pickle.dumps(b))):
        np.testing.assert_array_equal(c.cov_useritems, b.cov_useritems)
        assert c.kl_divergence() == pytest.approx(b.kl_divergence(), rel=1e-12)


def test_mn_driver_two_steps(M, ref_drivers):
    """the reference's own mn_active_pmf.compare / full_test on the GPU-backed MNActivePMF"""
    import random
    np.random.seed(2); random.seed(2)
    make_fake_data = ref_drivers("active_pmf").make_fake_data
    drv = ref_drivers("mn_active_pmf")
    assert drv.MNActivePMF is M.MNActivePMF
    real, ratings, vals = make_fake_data(noise=.25, num_users=6, num_items=6, rank=2,
                                         data_type='binary', mask_type='diag')
    res = drv.compare(['pred-variance', 'prob-ge-.5'], real, ratings, rating_vals=vals, latent_d=2,
                    steps=3, discrete_exp=True, do_threading=True)
    for k in ('pred-variance', 'prob-ge-.5'):
        assert len(res[k]) == 3 and res[k][2][0] == len(ratings) + 2
        assert res[k][1][4].shape == (6, 6)


def test_mn_wide_path_matches_batched_kernel(M, golden):
    """the host-driven wide fit (cuSOLVER dense algebra + sparse-mode kernels) walks the same
    line search as the one-CTA kernel and the reference"""
    g = golden("matrix_normal")
    b = M.MNActivePMF(g["b_ratings"], 2, rating_values={0, 1}, discrete_expectations=True)
    b.users, b.items = g["b_users"].copy(), g["b_items"].copy()
    b.wide_threshold = 0
    b.initialize_approx()
    kls = list(b.fit_normal_kls())
    assert len(kls) == len(g["b_kls"])
    np.testing.assert_allclose(kls, g["b_kls"], rtol=1e-8)
    np.testing.assert_allclose(b.cov_useritems, g["b_sig"], rtol=1e-6, atol=1e-8)
    pool = list(zip(g["b_cand_i"].tolist(), g["b_cand_j"].tolist()))
    b.mean, b.cov_useritems, b.cov_latents = g["b_mean"].copy(), g["b_sig"].copy(), g["b_om"].copy()
    ent = np.array(b._get_key_vals(pool[:2], M.MNActivePMF.exp_approx_entropy, None, None))
    np.testing.assert_allclose(ent, g["b_uv_entropy"][:2], rtol=1e-5)
