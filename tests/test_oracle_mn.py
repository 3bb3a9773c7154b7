"""Pins the matrix-normal part of the oracle against the reference's own
Cython build (tests/golden/matrix_normal.npz).  CPU only."""
import numpy as np
import pytest

from oracle import pmf_oracle as O


def test_mn_kl_gradient_criteria(golden):
    g = golden("matrix_normal")
    R, mean, sig, om = g["a_ratings"], g["a_mean"], g["a_sig"], g["a_om"]
    h = dict(zip(("sigma_sq", "sigma_u_sq", "sigma_v_sq"), g["a_hyp"]))
    nu = g["a_users"].shape[0]
    assert O.mn_kl_divergence(nu, R, mean, sig, om, **h) == pytest.approx(float(g["a_kl"]), rel=1e-12)
    gm, gs, go = O.mn_gradient(nu, R, mean, sig, om, **h)
    np.testing.assert_allclose(gm, g["a_gm"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(gs, g["a_gs"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(go, g["a_go"], rtol=1e-10, atol=1e-10)
    mv = np.array([O.mn_pred_mean_var(nu, mean, sig, om, i, j)
                   for i, j in zip(g["a_cand_i"], g["a_cand_j"])])
    np.testing.assert_allclose(mv[:, 0], g["a_pred_mean"], rtol=1e-12)
    np.testing.assert_allclose(mv[:, 1], g["a_pred_var"], rtol=1e-9)
    np.testing.assert_allclose(O.prob_ge_cutoff(mv[:, 0], mv[:, 1], 3.5), g["a_prob_ge_3_5"],
                               rtol=1e-9, atol=1e-300)
    assert O.mn_entropy(sig, om) == pytest.approx(float(g["a_entropy"]), rel=1e-12)


def test_mn_fit_trajectory(golden):
    g = golden("matrix_normal")
    R, U, V = g["b_ratings"], g["b_users"], g["b_items"]
    nu, d = U.shape
    mean0 = np.vstack((U, V))
    mean, sig, om, kls = O.mn_fit_normal_kls(nu, R, mean0, np.eye(mean0.shape[0]), np.eye(d))
    assert len(kls) == len(g["b_kls"])
    np.testing.assert_allclose(kls, g["b_kls"], rtol=1e-9)
    np.testing.assert_allclose(sig, g["b_sig"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(om, g["b_om"], rtol=1e-6, atol=1e-9)
    assert O.mn_entropy(sig, om) == pytest.approx(float(g["b_entropy"]), rel=1e-8)
