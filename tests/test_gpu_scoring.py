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


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-5)])
@pytest.mark.parametrize("n,m,d,nc,tile_bytes", [(3000, 2000, 32, 300_000, 64 * 1024),
                                                  (500, 777, 10, 50_000, 4096),
                                                  (64, 50, 5, 3000, 64 * 1024),
                                                  (200, 1500, 30, 20_000, 16 * 1024),
                                                  (20, 300, 7, 5000, 64 * 1024),     # runs longer than a segment
                                                  (50, 9000, 32, 30_000, 512),      # 2250 tiles: more than a CTA has threads
                                                  (300, 9000, 32, 400_000, 64 * 1024),
                                                  (30, 20, 2, 600, 1024)])
def test_tiled_pool_matches_flat_and_oracle(S, n, m, d, nc, tile_bytes, dtype, tol):
    """the shared-memory-tiled pool kernel (TMA-loaded item tiles) gives the same scores, in the
    caller's order, and the same winner as the flat kernel and the oracle"""
    import torch
    from active_matrix_factorization_b200 import device as D
    rng = np.random.RandomState(n + d)
    U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    ii, jj = rng.randint(0, n, nc), rng.randint(0, m, nc)       # caller's order: unsorted
    ii[5], jj[5] = ii[nc - 7], jj[nc - 7]                        # a duplicate pair -> exact tie
    flat, (fv, fi) = S.score_pred(U, V, ii, jj, dtype)
    pool = S.Pool(ii, jj, n, m, dtype, d, tile_bytes=tile_bytes)
    Ut, Vt = pool.pad(U), pool.pad(V)
    for maximize in (True, False):
        sc, best = pool.score_pred(Ut, Vt, want_scores=True, maximize=maximize)
        got = sc.double().cpu().numpy()
        bv, bi = S.unpack_best(best)
        np.testing.assert_allclose(got, flat, rtol=tol, atol=tol * np.abs(flat).max())
        ref = np.einsum("nd,nd->n", U[ii], V[jj])
        np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * np.abs(ref).max())
        want = int(np.argmax(got) if maximize else np.argmin(got))
        assert bi == want and bv == got[want]
    _, best = pool.score_pred(Ut, Vt, want_scores=False, index_base=1000)
    assert S.unpack_best(best)[1] == fi + 1000
    pool.close()


def test_tiled_pool_rejects_rows_wider_than_256_bytes(S):
    with pytest.raises(ValueError):
        S.Pool(np.array([0]), np.array([0]), 4, 4, "f64", 48)


def test_tiled_pool_ties_and_empty(S):
    from active_matrix_factorization_b200 import device as D
    U = np.ones((5, 3)); V = np.ones((40, 3))
    ii = np.array([4, 1, 2, 3, 0, 0]); jj = np.array([39, 1, 20, 3, 0, 17])
    pool = S.Pool(ii, jj, 5, 40, "f64", 3, tile_bytes=512)   # several tiles, all tied
    Ut, Vt = pool.pad(U), pool.pad(V)
    _, best = pool.score_pred(Ut, Vt)
    assert S.unpack_best(best) == (3.0, 0)                       # first in the caller's order
    empty = S.Pool(ii[:0], jj[:0], 5, 40, "f64", 3)
    _, best = empty.score_pred(Ut, Vt)
    assert S.unpack_best(best)[1] == -1


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
def test_block_diagonal_view_matches_oracle(S, dtype, tol):
    """Per-row posterior blocks (cov_uv = NULL): the criterion of candidate (i, j) touches only
    A_i, B_j and the two mean rows, so it is checked at any scale by giving the oracle a tiny
    2-row model that holds those blocks (hard parts)."""
    import torch
    from active_matrix_factorization_b200 import _native as N
    from active_matrix_factorization_b200 import device as D
    rng = np.random.RandomState(3)
    n, m, d, nc = 50, 40, 10, 2000
    mu, mv = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    def spd(count):
        x = rng.normal(size=(count, d, d))
        return np.einsum("bij,bkj->bik", x, x) / d + np.eye(d) * .1
    A, B = spd(n), spd(m)
    ii, jj = rng.randint(0, n, nc), rng.randint(0, m, nc)
    dt = D.np_dtype(dtype)
    t = {k: D.to_device(v, dt) for k, v in dict(mu=mu, mv=mv, A=A, B=B).items()}
    view = N.NormalView(t["mu"].data_ptr(), d, t["mv"].data_ptr(), d,
                        t["A"].data_ptr(), d * d, d, t["B"].data_ptr(), d * d, d, None, 0, 0, 0)
    ci, cj = D.to_device(ii, np.int32), D.to_device(jj, np.int32)
    for crit, cutoff in ((N.CRIT_APPROX_MEAN, 0.), (N.CRIT_PRED_VARIANCE, 0.), (N.CRIT_PROB_GE, 3.5)):
        sc, best = S.score_device(crit, dtype, ci, cj, d, view=view, cutoff=cutoff)
        got = sc.double().cpu().numpy()
        ref = np.empty(nc)
        for c in range(0, nc, 7):        # the oracle on a 1-user, 1-item model holding the blocks
            u, v = O.index_maps(1, 1, d)
            mean = np.concatenate((mu[ii[c]], mv[jj[c]]))
            cov = np.zeros((2 * d, 2 * d)); cov[:d, :d] = A[ii[c]]; cov[d:, d:] = B[jj[c]]
            e, var = O.pred_mean_var(u, v, mean, cov, 0, 0)
            ref[c] = {N.CRIT_APPROX_MEAN: e, N.CRIT_PRED_VARIANCE: var}.get(crit, O.prob_ge_cutoff(e, var, 3.5))
        sel = np.arange(0, nc, 7)
        ptol = tol if crit != N.CRIT_PROB_GE else tol * 50
        np.testing.assert_allclose(got[sel], ref[sel], rtol=ptol, atol=ptol * np.abs(ref[sel]).max())
        assert S.unpack_best(best)[1] == int(np.argmax(got))


def test_pool_remove_matches_shrinking_set(S):
    """an active loop on a device-resident pool: query the winner, remove it, score again --
    the winners come out in the order of the sorted scores, like max() over a shrinking set"""
    rng = np.random.RandomState(11)
    n, m, d, nc = 300, 700, 32, 40_000
    U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    ii, jj = rng.randint(0, n, nc), rng.randint(0, m, nc)
    pool = S.Pool(ii, jj, n, m, "f64", d)
    Ut, Vt = pool.pad(U), pool.pad(V)
    ref = np.einsum("nd,nd->n", U[ii], V[jj])
    order = np.argsort(-ref, kind="stable")
    picked = []
    for step in range(6):
        _, best = pool.score_pred(Ut, Vt)
        v, idx = S.unpack_best(best)
        assert v == pytest.approx(ref[idx], rel=1e-12)
        picked.append(idx)
        pool.remove([idx])
    assert picked == order[:6].tolist()
    pool.remove(order[6:20])                       # batch removal
    _, best = pool.score_pred(Ut, Vt)
    assert S.unpack_best(best)[1] == int(order[20])
    sc, _ = pool.score_pred(Ut, Vt, want_scores=True)
    alive = np.ones(nc, bool); alive[order[:20]] = False
    np.testing.assert_allclose(sc.cpu().numpy()[alive], ref[alive], rtol=1e-11, atol=1e-12)
    pool.close()


def test_best_reduce_matches_host_rule(S):
    """amf_best_reduce (the reduction of all-gathered per-GPU winners) applies the rule of
    parallel.reduce_winners: best value, lowest index on ties, index < 0 and NaN never win"""
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D, parallel as P
    lib = N.require_device()
    cases = [([1.0, 3.0, 3.0, 2.0], [7, 9, 4, 1]),
             ([float("nan"), -1.0, 5.0, 5.0], [0, 3, -1, 8]),
             ([float("nan")], [2]),
             ([2.5, 2.5], [-1, -1])]
    for vals, idx in cases:
        rec = torch.empty((len(vals), 2), dtype=torch.int64, device="cuda")
        rec[:, 0] = torch.tensor(vals, dtype=torch.float64, device="cuda").view(torch.int64)
        rec[:, 1] = torch.tensor(idx, dtype=torch.int64, device="cuda")
        for maximize in (True, False):
            out = torch.zeros(2, dtype=torch.int64, device="cuda")
            N.check(lib.amf_best_reduce(D.ptr(rec), len(vals), 1 if maximize else 0, D.ptr(out),
                                        D.stream_ptr()))
            v, i = S.unpack_best(out)
            hv, hi = P.reduce_winners(torch.tensor(vals, dtype=torch.float64), torch.tensor(idx), maximize)
            assert i == hi
            if i >= 0:
                assert v == hv


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-5)])
def test_host_csr_pool_matches_pairs(S, dtype, tol):
    """amf_score_pred_host_csr: a pool given as row offsets scores like the (i, j) form; users
    without candidates and an empty pool are fine"""
    rng = np.random.RandomState(4)
    n, m, d, nc = 300, 500, 10, 20_000
    U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    ii = np.sort(rng.randint(0, n // 2, nc) * 2)            # odd users have no candidates
    jj = rng.randint(0, m, nc)
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, ii + 1, 1)
    ptr = np.cumsum(ptr)
    ref = np.einsum("nd,nd->n", U[ii], V[jj])
    for maximize in (True, False):
        sc, (bv, bi) = S.score_pred_host_csr(U, V, ptr, jj, dtype, want_scores=True, maximize=maximize)
        np.testing.assert_allclose(sc, ref, rtol=tol, atol=tol * np.abs(ref).max())
        assert bi == int(np.argmax(sc) if maximize else np.argmin(sc)) and bv == sc[bi]
    _, (bv, bi) = S.score_pred_host_csr(U, V, ptr, jj, dtype)
    assert bi == int(np.argmax(sc if False else ref.astype(sc.dtype))) or abs(ref[bi] - ref.max()) <= tol * abs(ref.max())
    # 16-bit item ids (amf_score_pred_host_csr16): same scores and winner, bit for bit
    sc_min = sc
    sc16, best16 = S.score_pred_host_csr(U, V, ptr, jj.astype(np.uint16), dtype, want_scores=True,
                                         maximize=False)
    assert np.array_equal(sc16, sc_min) and best16 == (sc_min[int(np.argmin(sc_min))], int(np.argmin(sc_min)))
    with pytest.raises(RuntimeError):      # more than 65536 items do not fit
        S.score_pred_host_csr(U, np.zeros((70000, d)), ptr, jj.astype(np.uint16), dtype)
    _, (bv, bi) = S.score_pred_host_csr(U, V, np.zeros(n + 1, np.int64), jj[:0], dtype)
    assert bi == -1
    with pytest.raises(ValueError):
        S.score_pred_host_csr(U, V, ptr[:-1], jj, dtype)
