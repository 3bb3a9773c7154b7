"""GPU parity of candidate scoring + fused arg-best against the reference's golden outputs."""
import numpy as np
import pytest

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import scoring
    return scoring


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-5)])
@pytest.mark.parametrize("name", ["known_answer_10x10_d2", "random_12x20_d5"])
def test_criteria_golden(S, golden, name, dtype, tol):
    from active_matrix_factorization_b200 import _native as N
    g = golden(name)
    U, V = g["users"], g["items"]
    n, d = U.shape
    m = V.shape[0]
    ii, jj = g["cand_i"], g["cand_j"]
    pred, (bv, bi) = S.score_pred(U, V, ii, jj, dtype)
    np.testing.assert_allclose(pred, g["pred"], rtol=tol, atol=tol * np.abs(g["pred"]).max())
    assert bi == int(np.argmax(pred)) and bv == pytest.approx(pred.max())
    _, (bv, bi) = S.score_pred(U, V, ii, jj, dtype, maximize=False)
    assert bi == int(np.argmin(pred))
    for crit, key, cutoff in ((N.CRIT_APPROX_MEAN, "pred_mean", 0.), (N.CRIT_PRED_VARIANCE, "pred_variance", 0.),
                              (N.CRIT_PROB_GE, "prob_ge_half", .5), (N.CRIT_PROB_GE, "prob_ge_3_5", 3.5)):
        vals, (bv, bi) = S.score_normal(crit, g["mean"], g["cov"], n, m, d, ii, jj, dtype, cutoff=cutoff)
        scale = np.abs(g[key]).max()
        ptol = tol if crit != N.CRIT_PROB_GE else max(tol, 1e-9) * 50   # erfc amplifies input error
        np.testing.assert_allclose(vals, g[key], rtol=ptol, atol=ptol * scale, err_msg=key)
        assert vals[bi] == vals.max() and bi == int(np.argmax(vals))


def test_argbest_ties_and_empty(S):
    U = np.ones((5, 3)); V = np.ones((4, 3))
    ii = np.array([0, 1, 2, 3, 4, 0]); jj = np.array([0, 1, 2, 3, 0, 1])
    vals, (bv, bi) = S.score_pred(U, V, ii, jj, "f64")
    assert bi == 0 and bv == 3.0                         # all tied: the first candidate wins
    vals, (bv, bi) = S.score_pred(U, V, ii[:0], jj[:0], "f64")
    assert bi == -1 and len(vals) == 0                   # empty pool


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-5)])
def test_pred_large_random(S, dtype, tol):
    rng = np.random.RandomState(5)
    n, m, d, nc = 3000, 2000, 32, 500_000
    U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    ii, jj = np.sort(rng.randint(0, n, nc)), rng.randint(0, m, nc)
    vals, (bv, bi) = S.score_pred(U, V, ii, jj, dtype)
    ref = np.einsum("nd,nd->n", U[ii], V[jj])
    np.testing.assert_allclose(vals, ref, rtol=tol, atol=tol * np.abs(ref).max())
    assert vals[bi] == vals.max() and bi == int(np.argmax(vals))
