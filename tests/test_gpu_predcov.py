"""GPU kernels behind approx_pred_covs / _pred_entropy_bound (csrc/predcov.cu) against numpy: the
Isserlis covariance of all pairs of predicted cells (checked against a Monte-Carlo estimate-free
closed form evaluated with einsum, and against sampling for one entry) and slogdet by LU."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from active_matrix_factorization_b200 import build
    build.build()
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D, active_pmf as A
    return N, D, A, torch


def einsum_pred_covs(mean, cov, n, m, d):
    """the six Isserlis terms of active_pmf.py:324-390 with numpy einsum (fp64)"""
    nu = n * d
    mu, mv = mean[:nu].reshape(n, d), mean[nu:].reshape(m, d)
    Suu = cov[:nu, :nu].reshape(n, d, n, d)
    Svv = cov[nu:, nu:].reshape(m, d, m, d)
    Suv = cov[:nu, nu:].reshape(n, d, m, d)
    e = np.einsum
    out = (e('ik,al,jkbl->ijab', mu, mu, Svv) + e('ik,bl,aljk->ijab', mu, mv, Suv)
           + e('jk,al,ikbl->ijab', mv, mu, Suv) + e('jk,bl,ikal->ijab', mv, mv, Suu)
           + e('ikal,jkbl->ijab', Suu, Svv) + e('ikbl,aljk->ijab', Suv, Suv))
    return out.reshape(n * m, n * m)


@pytest.mark.parametrize("n,m,d,B", [(3, 4, 2, 3), (6, 7, 2, 2), (4, 3, 5, 1)])
def test_pred_covs_kernel_matches_the_isserlis_sum(env, n, m, d, B):
    N, D, A, torch = env
    rng = np.random.RandomState(n * m + d)
    k = (n + m) * d
    means = rng.normal(size=(B, k))
    L = rng.normal(size=(B, k, k)) / np.sqrt(k)
    covs = L @ L.transpose(0, 2, 1) + 0.1 * np.eye(k)
    got = A._pred_covs_device(D.to_device(means, np.float64), D.to_device(covs, np.float64), n, m, d).cpu().numpy()
    for b in range(B):
        want = einsum_pred_covs(means[b], covs[b], n, m, d)
        np.testing.assert_allclose(got[b], want, rtol=1e-12, atol=1e-13 * np.abs(want).max())
        np.testing.assert_allclose(got[b], got[b].T, rtol=1e-12, atol=1e-13 * np.abs(want).max())
    # the single-problem host wrapper and the diagonal = pred_variance's closed form
    one = A._pred_covs(means[0], covs[0], n, m, d)
    np.testing.assert_allclose(one, got[0], rtol=0, atol=0)
    # Monte-Carlo sanity of one off-diagonal entry
    X = rng.multivariate_normal(means[0], covs[0], size=200000)
    U, V = X[:, :n * d].reshape(-1, n, d), X[:, n * d:].reshape(-1, m, d)
    p01 = (U[:, 0] * V[:, 1]).sum(1)
    p12 = (U[:, 1] * V[:, 2]).sum(1)
    mc = np.cov(p01, p12)[0, 1]
    assert abs(mc - got[0][0 * m + 1, 1 * m + 2]) < 0.05 * np.sqrt(got[0][1, 1] * got[0][m + 2, m + 2]) + 0.02


@pytest.mark.parametrize("k", [1, 5, 42, 100, 257])
def test_slogdet_kernel_matches_numpy(env, k):
    N, D, A, torch = env
    rng = np.random.RandomState(k)
    mats = rng.normal(size=(6, k, k))
    mats[1] = mats[1] @ mats[1].T + np.eye(k)                 # positive definite
    mats[2][[0, k - 1]] = mats[2][[k - 1, 0]]                 # a row swap flips the sign (k > 1)
    mats[3] = mats[3] * 1e-3                                  # large negative log det
    if k > 1:
        mats[4][1] = 2 * mats[4][0]                           # singular
    want = [np.linalg.slogdet(a) for a in mats]
    sign, logdet = A._slogdet_device(D.to_device(mats, np.float64))
    for b, (ws, wl) in enumerate(want):
        if k > 1 and b == 4:
            assert sign[b] == 0 or logdet[b] < -25
            continue
        assert sign[b] == ws
        assert logdet[b] == pytest.approx(wl, rel=1e-10, abs=1e-9)
    s1, l1 = A._slogdet(mats[1])
    assert s1 == 1.0 and l1 == pytest.approx(want[1][1], rel=1e-12)
