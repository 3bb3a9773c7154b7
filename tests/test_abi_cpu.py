"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, and
the host-side mirror of the reference API behaves like the reference where no compute is
involved (constructors, bookkeeping, errors, pickling)."""
import copy
import os
import pickle
import re

import numpy as np
import pytest

from active_matrix_factorization_b200 import _native as N
from active_matrix_factorization_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "amf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "libamf_b200.so does not export %s" % name
    # and the ctypes prototypes cover the same set
    assert set(names) == set(N.PROTOTYPES) | {"amf_last_error"}


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        N.require_device()


def _toy():
    rng = np.random.RandomState(0)
    return np.array([(i, j, float(rng.randint(1, 6))) for i in range(4) for j in range(5)
                     if (i + j) % 2 == 0])


def test_constructor_and_bookkeeping():
    from active_matrix_factorization_b200.pmf_cy import ProbabilisticMatrixFactorization as PMF
    R = _toy()
    np.random.seed(3)
    p = PMF(R, 3)
    np.random.seed(3)
    assert np.array_equal(p.users, np.random.random((4, 3)))       # same RNG draws as the reference
    assert np.array_equal(p.items, np.random.random((5, 3)))
    assert (p.num_users, p.num_items, p.latent_d) == (4, 5, 3)
    assert p.mean_rating == pytest.approx(R[:, 2].mean())
    assert len(p.rated) == len(R) and len(p.rated) + len(p.unrated) == 20
    assert (p.sigma_sq, p.sigma_u_sq, p.sigma_v_sq) == (1, 10, 10)
    assert (p.learning_rate, p.min_learning_rate, p.stop_thresh) == (1e-4, 1e-10, 1e-2)
    p.add_rating(0, 1, 2.0)
    assert (0, 1) in p.rated and (0, 1) not in p.unrated and p.ratings.shape == (len(R) + 1, 3)
    with pytest.raises(ValueError):
        p.add_rating(0, 1, 3.0)
    with pytest.raises(TypeError):
        p.add_ratings([[1, 2]])
    with pytest.raises(AssertionError):
        p.add_rating(9, 0, 1.0)
    with pytest.raises(TypeError):
        PMF(np.zeros((3, 2)))
    with pytest.raises(TypeError):
        PMF(None)
    p.fit_type = ('nope',)
    with pytest.raises(ValueError):
        p.do_fit()
    q = PMF(R, 2, knowable=[(0, 1), (0, 0)])
    assert q.unrated == {(0, 1)}


def test_state_roundtrip():
    from active_matrix_factorization_b200.pmf_cy import ProbabilisticMatrixFactorization as PMF
    p = PMF(_toy(), 2, True)
    p.sigma_sq = 0.5
    keys = {'latent_d', 'num_users', 'num_items', 'learning_rate', 'min_learning_rate',
            'stop_thresh', 'sigma_sq', 'sigma_u_sq', 'sigma_v_sq', 'fit_type', 'sig_u_mean',
            'sig_u_var', 'sig_v_mean', 'sig_v_var', 'ratings', 'users', 'items', 'subtract_mean',
            'mean_rating', 'rated', 'unrated'}
    assert set(p.__getstate__()) == keys                               # pmf_cy.pyx:99-126
    for q in (copy.deepcopy(p), pickle.loads(pickle.dumps(p))):
        assert q.sigma_sq == 0.5 and q.subtract_mean and q.latent_d == 2
        assert np.array_equal(q.users, p.users) and q.users is not p.users
        assert q.rated == p.rated
    r = PMF(_toy(), 1)
    r.__setstate__(p.__getstate__())                                   # add_rmse_boosts.py:39-40
    assert r.latent_d == 2 and np.array_equal(r.items, p.items)


def test_parse_fit_type():
    from active_matrix_factorization_b200.pmf_cy import parse_fit_type
    assert parse_fit_type('batch') == ('batch',)
    assert parse_fit_type('mini-valid,100,30,1.5') == ('mini-valid', 100, 30, 1.5)


def test_error_conventions_of_the_active_classes():
    """SURVEY.md 8b error conventions that need no device: same exception types and messages
    as the reference (active_pmf.py:121-122,206-207,731-732; pmf_cy.pyx:144-145;
    normal_exps_cy.pyx:149-150; bayes_pmf.py:88-89)."""
    from active_matrix_factorization_b200 import active_pmf as A, bayes_pmf as Bm, mn_active_pmf as M
    R = _toy()
    with pytest.raises(ValueError, match="got ratings not in rating_values"):
        A.ActivePMF(R, 2, rating_values={7, 8})
    with pytest.raises(ValueError, match="got ratings not in rating_values"):
        Bm.BayesianPMF(R, 2, rating_values={7, 8})
    a = A.ActivePMF(R, 2, rating_values={1, 2, 3, 4, 5})
    assert a.rating_values == (1., 2., 3., 4., 5.) and a.rating_bounds[0] == -np.inf
    assert a.rating_bounds[1] == 1.5 and a.rating_bounds[-1] == np.inf
    with pytest.raises(ValueError, match="got ratings with bad values"):
        a.add_rating(0, 1, 9.0)
    with pytest.raises(ValueError, match="run initialize_approx first"):
        a.kl_divergence()
    with pytest.raises(TypeError, match="run initialize_approx first"):
        A.normal_gradient(a)
    with pytest.raises(ValueError, match="empty pool"):
        a.pick_query_point(pool=[])
    assert a.pick_query_point(pool=[(3, 3)]) == (3, 3)           # single element: no evaluation
    assert a.approx_dim == (4 + 5) * 2 and a.u.shape == (2, 4) and a.v[0, 0] == 8
    mn = M.MNActivePMF(R, 2)
    with pytest.raises(ValueError, match="run initialize_approx first"):
        mn.kl_divergence()
    with pytest.raises(TypeError):
        M.matrixnormal_gradient(mn)
    assert not hasattr(mn, "cov") and mn.cov_useritems is None
    assert set(M.KEY_FUNCS) == set(A.KEY_FUNCS) - {"pred-entropy-bound", "pred-entropy-bound-approx"}
    for f in A.KEY_FUNCS.values():
        assert f.chooser in (min, max)
    b = Bm.BayesianPMF(R, 3)
    assert b.beta == 2 and b.subtract_mean and b.u_hyperparams[2] == 3
    st = b.__getstate__()
    assert {'discrete_expectations', 'rating_values', 'beta', 'u_hyperparams', 'v_hyperparams'} <= set(st)


def test_pool_arrays_conversion():
    """host logic: a list of (i, j) pairs becomes two int32 arrays in iteration order whatever
    the pair type (tuples, numpy integers, lists of floats holding integers)"""
    from active_matrix_factorization_b200 import active_pmf as A
    ii, jj = A._pool_arrays([(1, 2), (3, 4), (0, 9)])
    assert ii.dtype == np.int32 and ii.tolist() == [1, 3, 0] and jj.tolist() == [2, 4, 9]
    ii, jj = A._pool_arrays([(np.int64(5), np.int32(6))])
    assert (ii.tolist(), jj.tolist()) == ([5], [6])
    ii, jj = A._pool_arrays([[7.0, 8.0], [1.0, 0.0]])
    assert (ii.tolist(), jj.tolist()) == ([7, 1], [8, 0])
    s = {(2, 3), (4, 5), (6, 7)}
    ii, jj = A._pool_arrays(list(s))
    assert list(zip(ii.tolist(), jj.tolist())) == list(s)


def test_ratings_append_bookkeeping_without_device():
    """host logic of add_ratings: duplicate cells and bad values are rejected before any device
    work, rated / unrated sets and the mean follow the reference (pmf_cy.pyx:128-156)"""
    from active_matrix_factorization_b200 import pmf_cy as P
    R = np.array([[0, 0, 1.], [1, 2, 3.], [2, 1, 5.]])
    p = P.ProbabilisticMatrixFactorization(R, 2)
    with pytest.raises(ValueError, match="already rated"):
        p.add_rating(1, 2, 4.)
    with pytest.raises(TypeError):
        p.add_ratings([[0, 1]])
    p.add_ratings([[0, 1, 2.], [2, 2, 4.]])        # no device handle exists yet: host arrays only
    assert p.ratings.shape == (5, 3) and (0, 1) in p.rated and (0, 1) not in p.unrated
    assert p.mean_rating == pytest.approx(3.0)


def test_shims_resolve_to_the_mirror():
    """drop-in boundary: with shims/ first on sys.path the reference's own import
    lines (`from pmf_cy import ...`, `import active_pmf`, ...) get the GPU-backed classes"""
    import importlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from pmf_cy import ProbabilisticMatrixFactorization, rmse, parse_fit_type\n"
        "import active_pmf, bayes_pmf, mn_active_pmf, normal_exps_cy, matrix_normal_exps_cy\n"
        "import active_matrix_factorization_b200.pmf_cy as M\n"
        "assert ProbabilisticMatrixFactorization is M.ProbabilisticMatrixFactorization\n"
        "assert active_pmf.ActivePMF.__module__.startswith('active_matrix_factorization_b200')\n"
        "assert len(active_pmf.KEY_FUNCS) == 15 and hasattr(bayes_pmf, 'BayesianPMF')\n"
        "assert hasattr(mn_active_pmf, 'MNActivePMF') and hasattr(normal_exps_cy, 'exp_dotprod_sq')\n"
        "print('ok')\n" % (root, os.path.join(root, "shims")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_which_indices_fast_paths_equal_numpy_indexing():
    """BayesianPMF._which_indices (host side of bayes_pmf.py:433-455's `[which]`): the index-pair
    and mask fast paths give exactly what indexing the n x m grids gives, in the same order"""
    from active_matrix_factorization_b200 import bayes_pmf
    rng = np.random.RandomState(0)
    n, m = 7, 9
    R = np.column_stack((rng.randint(0, n, 20), rng.randint(0, m, 20), rng.randint(1, 6, 20))).astype(float)
    R[0, :2] = (n - 1, m - 1)
    b = bayes_pmf.BayesianPMF(R, 2)
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
    mask = rng.uniform(size=(n, m)) < .4
    cases = [Ellipsis, None, mask, np.nonzero(mask), (np.array([0, 6, 3]), np.array([8, 0, 4])),
             (np.array([[0, 1], [2, 3]]), np.array([[4, 5], [6, 7]])),      # 2-D index arrays
             (np.array([-1, 2]), np.array([0, -2])),                        # negative ids wrap
             (slice(1, 4), np.array([0, 2])), ([1, 2], [3, 4]),
             (np.array([], dtype=int), np.array([], dtype=int))]
    for which in cases:
        w = Ellipsis if which is None else which
        gi, gj, shape = b._which_indices(which)
        assert shape == ii[w].shape
        np.testing.assert_array_equal(gi, ii[w].reshape(-1))
        np.testing.assert_array_equal(gj, jj[w].reshape(-1))
    with pytest.raises(IndexError):
        b._which_indices((np.array([n]), np.array([0])))


def test_load_coo_formats(tmp_path):
    """pmf_cy.load_coo: the same rating list through every accepted file form"""
    import pickle
    from active_matrix_factorization_b200 import pmf_cy
    rng = np.random.RandomState(1)
    i, j = rng.randint(0, 30, 200), rng.randint(0, 17, 200)
    i[0], j[0] = 29, 16
    r = rng.normal(size=200)
    table = np.column_stack((i, j, r)).astype(float)
    real = np.zeros((40, 20))
    np.savez(tmp_path / "a.npz", i=i.astype(np.int64), j=j.astype(np.int16), r=r.astype(np.float32))
    np.savez(tmp_path / "b.npz", i=i, j=j, r=r, shape=np.array([40, 20]))
    np.savez_compressed(tmp_path / "c.npz", _ratings=table, _real=real, _rating_vals=np.array([1, 2]))
    with open(tmp_path / "d.pkl", "wb") as f:
        pickle.dump({"_ratings": table, "_real": None}, f)
    np.save(tmp_path / "e.npy", table)
    np.savetxt(tmp_path / "f.txt", table)
    want = {"a.npz": (30, 17), "b.npz": (40, 20), "c.npz": (40, 20), "d.pkl": (30, 17),
            "e.npy": (30, 17), "f.txt": (30, 17)}
    for name, shape in want.items():
        gi, gj, gr, n, m = pmf_cy.load_coo(str(tmp_path / name))
        assert (n, m) == shape, name
        assert gi.dtype == np.int32 and gj.dtype == np.int32
        np.testing.assert_array_equal(gi, i)
        np.testing.assert_array_equal(gj, j)
        np.testing.assert_allclose(gr, r, rtol=1e-6 if name == "a.npz" else 1e-15)
    np.savez(tmp_path / "bad.npz", i=i, j=j, r=r, shape=np.array([10, 20]))
    with pytest.raises(ValueError):
        pmf_cy.load_coo(str(tmp_path / "bad.npz"))
    np.savez(tmp_path / "bad2.npz", x=i)
    with pytest.raises(ValueError):
        pmf_cy.load_coo(str(tmp_path / "bad2.npz"))
    np.save(tmp_path / "bad3.npy", table[:, :2])
    with pytest.raises(TypeError):
        pmf_cy.load_coo(str(tmp_path / "bad3.npy"))
