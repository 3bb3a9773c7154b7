"""The reference's own test for this path (python-pmf/test_normal_exps.py): Gaussian third and
fourth moments, pure-vs-Cython to 1e-7 and within 2 % of a Monte-Carlo estimate.  Same input
distributions here (tests/golden/moments.npz, make_golden.py moments, with the answers of the
compiled normal_exps_cy); checked against them are the oracle's restatement and the product's
host-side scalar formulas (no device needed for these).  CPU only."""
import numpy as np
import pytest

from oracle import pmf_oracle as O

SCALARS = (("tripexpect", 3, lambda x: x.prod(1)),
           ("quadexpect", 4, lambda x: x.prod(1)),
           ("exp_squared", 2, lambda x: (x ** 2).prod(1)),
           ("exp_a2bc", 3, lambda x: x[:, 0] ** 2 * x[:, 1] * x[:, 2]))


@pytest.fixture(scope="module")
def g(golden):
    return golden("moments")


def oracle_moment(name, mean, cov, idx):
    if name == "tripexpect":
        return O._trip(mean, cov, *idx)
    if name == "quadexpect":
        return O._e4(mean, cov, *idx)
    if name == "exp_squared":
        a, b = idx
        return O._e4(mean, cov, a, a, b, b)
    a, b, c = idx
    return O._e4(mean, cov, a, a, b, c)


@pytest.mark.parametrize("name,dim,monte", SCALARS)
def test_scalar_moments(g, name, dim, monte):
    from active_matrix_factorization_b200 import normal_exps_cy as mirror
    rng = np.random.RandomState(dim)
    for t in range(3):
        mean, cov, want = g["%s_mean%d" % (name, t)], g["%s_cov%d" % (name, t)], float(g["%s_val%d" % (name, t)])
        idx = tuple(range(dim))
        assert oracle_moment(name, mean, cov, idx) == pytest.approx(want, rel=1e-12)
        assert getattr(mirror, name)(mean, cov, *idx) == pytest.approx(want, rel=1e-12)
        samples = rng.multivariate_normal(mean, cov, 500_000)         # the reference's 2 % check
        assert monte(samples).mean() == pytest.approx(want, rel=.02)


def test_exp_dotprod_sq(g):
    n, m, d = 1, 1, 3
    u, v = O.index_maps(n, m, d)
    rng = np.random.RandomState(7)
    for t in range(3):
        mean, cov = g["exp_dotprod_sq_mean%d" % t], g["exp_dotprod_sq_cov%d" % t]
        want = float(g["exp_dotprod_sq_val%d" % t])
        assert O.exp_dotprod_sq(u, v, mean, cov, 0, 0) == pytest.approx(want, rel=1e-12)
        e, var = O.pred_mean_var(u, v, mean, cov, 0, 0)
        assert var + e * e == pytest.approx(want, rel=1e-10)           # closed form used on the GPU
        x = rng.multivariate_normal(mean, cov, 500_000)
        assert ((x[:, :3] * x[:, 3:]).sum(1) ** 2).mean() == pytest.approx(want, rel=.02)
