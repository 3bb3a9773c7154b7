"""GPU parity of the fits and helpers around the main path:
sigma-learning line search, momentum SGD with a validation split, dense prediction, RMSE and
bayes_rmse -- against tests/golden/extras_30x40_d4.npz (reference Cython, make_golden.py extras).
Parity mode (f64); the trajectories are sequences of accept/reject decisions, so they are
compared step for step."""
import random
from itertools import islice

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def amf():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import pmf_cy
    return pmf_cy


@pytest.fixture(scope="module")
def g(golden):
    return golden("extras_30x40_d4")


def model(amf, g, sm, scale=1.0, cls=None, **kw):
    p = (cls or amf.ProbabilisticMatrixFactorization)(g["ratings"].copy(), 4, sm, **kw)
    p.compute_dtype = "f64"
    p.users, p.items = g["users0"] * scale, g["items0"] * scale
    return p


@pytest.mark.parametrize("tag", ["", "_sm", "_prior"])
def test_fit_with_sigmas_trajectory(amf, g, tag):
    """pmf_cy.pyx:384-403: the variances change between accepted steps while fit_lls keeps
    comparing with the objective it yielded (:285)."""
    p = model(amf, g, tag == "_sm")
    if tag == "_prior":
        p.sig_u_mean, p.sig_u_var, p.sig_v_mean, p.sig_v_var = g["sig_prior"]
    ref = g["ws_lls" + tag]
    lls, sig = [], []
    for ll in islice(p.fit_with_sigmas_lls(5, 2), len(ref)):
        lls.append(ll)
        sig.append((p.sigma_sq, p.sigma_u_sq, p.sigma_v_sq))
    assert len(lls) == len(ref)
    np.testing.assert_allclose(lls[:150], ref[:150], rtol=1e-9)
    np.testing.assert_allclose(sig[:150], g["ws_sigmas" + tag][:150], rtol=1e-8)
    np.testing.assert_allclose(lls, ref, rtol=1e-5)
    np.testing.assert_allclose(p.users, g["ws_users" + tag], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(p.items, g["ws_items" + tag], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("sm", [False, True])
def test_minibatch_validation_and_dense_outputs(amf, g, sm):
    """pmf_cy.pyx:308-381 with seeded numpy / stdlib streams; lr / n is a float quotient."""
    tag = "_sm" if sm else ""
    p = model(amf, g, sm, .3)
    np.random.seed(3); random.seed(3)
    errs = list(islice(p.fit_minibatches_validation(50, 40, lr=.05), 6))
    np.testing.assert_allclose(errs, g["mb_errs" + tag], rtol=1e-6)
    np.testing.assert_allclose(p.users, g["mb_users" + tag], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(p.items, g["mb_items" + tag], rtol=1e-9, atol=1e-12)
    # dense prediction and the three forms of rmse (pmf_cy.pyx:410-426)
    np.testing.assert_allclose(p.predicted_matrix(), g["pm" + tag], rtol=1e-9, atol=1e-12)
    real = g["real"]
    got = [p.rmse(real), p.rmse(real, g["mask"]), p.rmse(real, g["rmse_rows"])]
    np.testing.assert_allclose(got, g["rmse3" + tag], rtol=1e-6)
    assert amf.rmse(p.predicted_matrix(), real) == pytest.approx(g["rmse3" + tag][0], rel=1e-6)

    q = model(amf, g, sm, .3)
    np.random.seed(3); random.seed(3)
    q.fit_minibatches_until_validation(50, 40, lr=.05, stop_thresh=1e-3)
    np.testing.assert_allclose(q.users, g["mbu_users" + tag], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(q.items, g["mbu_items" + tag], rtol=1e-9, atol=1e-12)
    # the 'mini-valid' fit type routes do_fit() there (pmf_cy.pyx:297-305)
    r = model(amf, g, sm, .3, fit_type=('mini-valid', 50, 40))
    r.users, r.items = g["users0"] * .3, g["items0"] * .3
    np.random.seed(3); random.seed(3)
    r.do_fit()
    assert np.isfinite(r.users).all() and not np.allclose(r.users, g["users0"] * .3)


def test_rmse_selection_forms_match_numpy_indexing(amf):
    """rmse(real, on) for every form of `on` the reference's real[on] accepts (pmf_cy.pyx:28-29,
    422-426): boolean mask, row indices, (rows, cols) index arrays -- also with a repeated cell,
    which real[on] counts twice -- on a matrix that is not a multiple of the kernel's 64 x 64 tile"""
    rng = np.random.RandomState(5)
    n, m, d = 70, 131, 5
    R = np.column_stack((rng.randint(0, n, 300), rng.randint(0, m, 300), rng.normal(3, 1, 300)))
    R[0, 0], R[1, 1] = n - 1, m - 1
    p = amf.ProbabilisticMatrixFactorization(R, d, subtract_mean=True)
    real = rng.normal(3, 1, (n, m))
    pred = p.users @ p.items.T + p.mean_rating
    np.testing.assert_allclose(p.predicted_matrix(), pred, rtol=1e-12, atol=1e-13)
    mask = rng.rand(n, m) < .3
    rows = np.array([3, 69, 0])
    cells = (rng.randint(0, n, 50), rng.randint(0, m, 50))
    twice = (np.array([1, 1, 2]), np.array([5, 5, 130]))
    for on in (None, mask, rows, cells, twice):
        want = np.sqrt(np.mean((real - pred) ** 2)) if on is None else np.sqrt(np.mean((real[on] - pred[on]) ** 2))
        assert p.rmse(real, on) == pytest.approx(float(np.float32(want)), rel=1e-6)


def test_bayes_rmse(amf, g):
    from active_matrix_factorization_b200 import bayes_pmf
    b = bayes_pmf.BayesianPMF(g["ratings"].copy(), 4)
    b.compute_dtype = "f64"
    samples = list(zip(g["br_samples_u"], g["br_samples_v"]))
    got = [b.bayes_rmse(samples, g["real"]), b.bayes_rmse(samples, g["real"], g["mask"])]
    np.testing.assert_allclose(got, g["bayes_rmse"], rtol=1e-6)
    # and the chain itself from the same seed (bayes_pmf.py:227-302)
    b.users, b.items = g["users0"].copy(), g["items0"].copy()
    np.random.seed(13)
    s = list(islice(b.samples(num_gibbs=2), 4))
    np.testing.assert_allclose(np.array([x[0] for x in s]), g["br_samples_u"], rtol=1e-6, atol=1e-8)
    np.random.seed(13)
    s2 = list(islice(b.samples_parallel(num_gibbs=2), 4))      # same chain, the launch is the fan-out
    np.testing.assert_allclose(s2[3][1], s[3][1], rtol=1e-12)
    with pytest.raises(ValueError):
        next(b.samples_parallel(multiproc_mode='force'))    # a generator, as in the reference
