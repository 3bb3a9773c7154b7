"""CPU tests of host-side logic added in round 2: the driver loader, the block-diagonal
container, the vectorised hyper-parameter draw, the rating-file loader's range checks."""
import pickle

import numpy as np
import pytest


def test_driver_loader_binds_reference_drivers_to_the_mirror_classes():
    """drivers.load executes the reference's own active_pmf / bayes_pmf / mn_active_pmf source
    (the patched copies under oracle/_ref) and rebinds classes and registries"""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref is not built")
    from active_matrix_factorization_b200 import active_pmf as A, bayes_pmf as Bm, drivers, mn_active_pmf as M
    a = drivers.load("active_pmf", ref_loader.REF_DIR)
    assert a.ActivePMF is A.ActivePMF and a.KEY_FUNCS is A.KEY_FUNCS
    for fn in ("full_test", "compare", "main", "make_fake_data", "get_ratings"):
        assert callable(getattr(a, fn)) and getattr(a, fn).__module__.startswith("amf_b200_reference_drivers")
        assert not hasattr(A, fn)                       # not restated in the package
    b = drivers.load("bayes_pmf", ref_loader.REF_DIR)
    assert b.BayesianPMF is Bm.BayesianPMF and b.KEYS is Bm.KEYS and callable(b.compare_active)
    m = drivers.load("mn_active_pmf", ref_loader.REF_DIR)
    assert m.MNActivePMF is M.MNActivePMF and callable(m.compare)
    # the reference's data generator runs as it is (same draws as the reference: same source)
    np.random.seed(0)
    real, ratings, vals = a.make_fake_data(noise=.25, num_users=5, num_items=4, rank=2,
                                           data_type='binary', mask_type='diag')
    assert real.shape == (5, 4) and set(vals) == {0, 1} and ratings.shape[1] == 3
    with pytest.raises(ValueError):
        drivers.load("pmf_cy", ref_loader.REF_DIR)


def test_in_process_pool_surface():
    from active_matrix_factorization_b200.drivers import InProcessPool
    p = InProcessPool(4)
    assert p.map(abs, [-1, 2]) == [1, 2] and p.apply(max, (1, 3)) == 3
    assert p.map_async(abs, [-2]).get() == [2] and p.apply_async(min, (4, 2)).get() == 2
    p.close(); p.join()


def test_block_diagonal_container():
    from active_matrix_factorization_b200.blocks import BlockDiagonal
    rng = np.random.RandomState(0)
    n, m, d = 3, 2, 2
    A = rng.normal(size=(n, d, d)); B = rng.normal(size=(m, d, d))
    bd = BlockDiagonal(rng.normal(size=(n, d)), rng.normal(size=(m, d)), A, B, A, B,
                       np.zeros((n, d)), np.zeros((m, d)), np.zeros(n), np.zeros(m))
    full = bd.toarray()
    assert bd.shape == (10, 10) and full.shape == (10, 10)
    np.testing.assert_array_equal(full[2:4, 2:4], A[1])
    np.testing.assert_array_equal(full[6:8, 6:8], B[0])
    assert full[0, 2] == 0 and bd.mean() == pytest.approx(full.mean())
    np.testing.assert_array_equal(np.asarray(bd), full)
    assert bd.stacked_mean().shape == (10,)
    again = pickle.loads(pickle.dumps(bd))
    np.testing.assert_array_equal(again.toarray(), full)
    with pytest.raises(MemoryError):
        bd.toarray(max_dim=4)


def test_batched_hyperparameter_draw_equals_the_scalar_one():
    """_hyperparams_batched with one chain consumes the numpy stream exactly like
    sample_hyperparam's d x d part (bayes_pmf.py:166-186): both Wishart branches and d = 1"""
    from active_matrix_factorization_b200 import bayes_pmf as Bm
    rng = np.random.RandomState(1)
    R = np.column_stack((rng.randint(0, 40, 300), rng.randint(0, 30, 300), rng.randint(1, 6, 300))).astype(float)
    R[0, :2] = (39, 29)
    for d, nrows in ((3, 15), (15, 943), (1, 10)):
        np.random.seed(2)
        b = Bm.BayesianPMF(R, d, knowable=())
        x = rng.normal(size=(nrows, d))
        xb = x.mean(0)
        sb = np.atleast_2d(np.cov(x.T)) if d > 1 else np.array(float(x.var(ddof=1)))
        np.random.seed(5)
        mu1, al1 = b._hyperparam_from_moments(nrows, xb, sb, True)
        np.random.seed(5)
        mu2, al2 = b._hyperparams_batched(nrows, xb[None], np.atleast_2d(sb)[None], True)
        np.testing.assert_allclose(mu2[0], mu1, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(al2[0], np.atleast_2d(al1), rtol=1e-13)
    # several chains at once: symmetric positive definite draws of the right shape
    P, d = 7, 4
    x_bar, s_bar = rng.normal(size=(P, d)), np.array([np.cov(rng.normal(size=(50, d)).T) for _ in range(P)])
    b = Bm.BayesianPMF(R, d, knowable=())
    mu, al = b._hyperparams_batched(50, x_bar, s_bar, False)
    assert mu.shape == (P, d) and al.shape == (P, d, d)
    assert np.allclose(al, al.transpose(0, 2, 1)) and np.all(np.linalg.eigvalsh(al) > 0)


def test_load_coo_checks_ids_before_narrowing(tmp_path):
    from active_matrix_factorization_b200.pmf_cy import load_coo
    p = str(tmp_path / "r.npz")
    np.savez(p, i=np.array([0, 2 ** 31 + 5], dtype=np.int64), j=np.array([1, 2], dtype=np.int64),
             r=np.array([1., 2.]))
    with pytest.raises(ValueError):                      # would wrap to a negative / small id as int32
        load_coo(p)
    np.savez(p, i=np.array([0, 3]), j=np.array([1, 2]), r=np.array([1., 2.]), shape=np.array([3, 3]))
    with pytest.raises(ValueError):
        load_coo(p)
    np.savez(p, i=np.array([0, 2]), j=np.array([1, 2]), r=np.array([1., 2.]))
    i, j, r, n, m = load_coo(p)
    assert (n, m) == (3, 3) and i.dtype == np.int32
