"""Active-loop residency of the rating list: amf_ratings_append keeps new ratings
as an unsorted tail on the device; the fused loss+gradient must see them at once, every consumer
of the sorted lists must see them after the (automatic) compaction, and the drop-in classes'
add_rating / add_ratings must give what a model built from the whole list gives.

Tolerances: f64 1e-10, f32 2e-5 relative (north_star: 1e-5 on the objective)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu

HYP = dict(sigma_sq=.8, sigma_u_sq=7., sigma_v_sq=12.)


@pytest.fixture(scope="module")
def env():
    from active_matrix_factorization_b200 import build
    build.build()
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    N.require_device()
    return N, D, torch


def rel_err(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def loss_grad(env, rat, U, V, d, dtype):
    N, D, torch = env
    lib = N.require_device()
    ld = D.padded_ld(d, dtype)
    Ut, Vt = D.to_padded(U, dtype), D.to_padded(V, dtype)
    dU, dV = torch.empty_like(Ut), torch.empty_like(Vt)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    params = D.pmf_params(HYP["sigma_sq"], HYP["sigma_u_sq"], HYP["sigma_v_sq"], 0.0)
    N.check(lib.amf_pmf_loss_grad(rat.handle, D.code(dtype), d, ld, D.ptr(Ut), D.ptr(Vt),
                                  C.byref(params), D.ptr(dU), D.ptr(dV), D.ptr(sums), D.stream_ptr()))
    s = sums.cpu().numpy()
    ll = -s[0] / (2 * HYP["sigma_sq"]) - s[1] / (2 * HYP["sigma_u_sq"]) - s[2] / (2 * HYP["sigma_v_sq"])
    return ll, dU[:, :d].double().cpu().numpy(), dV[:, :d].double().cpu().numpy()


@pytest.mark.parametrize("layout", ["rows", "tiled"])
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
def test_append_tail_then_compact(env, dtype, tol, layout):
    N, D, torch = env
    rng = np.random.RandomState(7)
    n, m, d, nnz = 900, 700, 32, 60000
    cells = rng.permutation(n * m)[:nnz]
    ii, jj, r = (cells // m).astype(np.int32), (cells % m).astype(np.int32), rng.normal(3, 1, nnz)
    R = np.column_stack((ii, jj, r)).astype(float)
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    k = 50000
    rat = D.Ratings(n, m, ii[:k], jj[:k], r[:k], dtype)
    rat.set_layout(layout)
    cuts = [k, k + 1, k + 4000, nnz]                      # a single rating, then two batches
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        rat.append(ii[lo:hi], jj[lo:hi], r[lo:hi])
        ll, gu, gv = loss_grad(env, rat, U, V, d, dtype)
        assert ll == pytest.approx(O.log_likelihood(R[:hi], U, V, **HYP), rel=tol)
        ou, ov = O.gradient(R[:hi], U, V, **HYP)
        assert rel_err(gu, ou) < tol and rel_err(gv, ov) < tol
    assert N.load().amf_ratings_nnz(rat.handle) == nnz
    rat.compact()
    ll2, gu2, gv2 = loss_grad(env, rat, U, V, d, dtype)
    assert ll2 == pytest.approx(ll, rel=tol) and rel_err(gu2, gu) < tol and rel_err(gv2, gv) < tol
    assert rat.mean() == pytest.approx(r.mean(), rel=1e-6 if dtype == "f32" else 1e-12)


def test_append_to_empty_and_auto_compaction(env):
    """a list created empty takes everything as tail; a tail above 65536 entries is merged by the
    append itself; the merged lists are those of a fresh handle (Gibbs conditionals bit-equal)"""
    N, D, torch = env
    lib = N.require_device()
    rng = np.random.RandomState(8)
    n, m, d, nnz = 400, 300, 6, 70000
    cells = rng.permutation(n * m)[:nnz]
    ii, jj, r = (cells // m).astype(np.int32), (cells % m).astype(np.int32), rng.normal(3, 1, nnz)
    R = np.column_stack((ii, jj, r)).astype(float)
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    rat = D.Ratings(n, m, ii[:0], jj[:0], r[:0], "f64")
    rat.append(ii[:100], jj[:100], r[:100])
    ll, gu, gv = loss_grad(env, rat, U, V, d, "f64")
    assert ll == pytest.approx(O.log_likelihood(R[:100], U, V, **HYP), rel=1e-10)
    rat.append(ii[100:], jj[100:], r[100:])                # > 65536: compacts
    ll, gu, gv = loss_grad(env, rat, U, V, d, "f64")
    ou, ov = O.gradient(R, U, V, **HYP)
    assert rel_err(gu, ou) < 1e-10 and rel_err(gv, ov) < 1e-10
    fresh = D.Ratings(n, m, ii, jj, r, "f64")
    alpha = torch.eye(d, dtype=torch.float64, device="cuda") * 2.0
    mu = torch.zeros(d, dtype=torch.float64, device="cuda")
    z = torch.from_numpy(rng.normal(size=(n, d))).cuda()
    Vt = torch.from_numpy(V).cuda()
    outs = []
    for h in (rat, fresh):
        out = torch.empty((n, d), dtype=torch.float64, device="cuda")
        N.check(lib.amf_gibbs_half_sweep(h.handle, 0, N.F64, d, D.ptr(Vt), D.ptr(alpha), D.ptr(mu), 2.0,
                                         0.0, D.ptr(z), D.ptr(out), D.stream_ptr()))
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("ctor", ["array", "coo"])
def test_model_add_ratings_keeps_device_list(env, ctor):
    """ProbabilisticMatrixFactorization.add_rating(s) (pmf_cy.pyx:128-156) on a model whose list
    already lives on the device: same objective, gradient and mean as a model built from
    everything, and the device handle is the same object (no re-upload)"""
    N, D, torch = env
    from active_matrix_factorization_b200.pmf_cy import ProbabilisticMatrixFactorization as PMF
    rng = np.random.RandomState(9)
    n, m, d, nnz = 120, 90, 5, 3000
    cells = rng.permutation(n * m)[:nnz]
    R = np.column_stack((cells // m, cells % m, rng.randint(1, 6, nnz))).astype(float)
    R[0, :2] = (n - 1, m - 1)
    U, V = rng.normal(0, .3, (n, d)), rng.normal(0, .3, (m, d))
    k = 2500

    def make(rows):
        if ctor == "array":
            p = PMF(rows, d)
            p.users, p.items = U.copy(), V.copy()
        else:
            p = PMF.from_coo(rows[:, 0].astype(np.int32), rows[:, 1].astype(np.int32), rows[:, 2], n, m, d,
                             init=(U.copy(), V.copy()))
        return p

    p = make(R[:k])
    p.log_likelihood()                                   # puts the list on the device
    handle = p._rating_handle()
    p.add_rating(*R[k])
    p.add_ratings(R[k + 1:])
    assert p._rating_handle() is handle
    full = make(R)
    assert p.mean_rating == pytest.approx(full.mean_rating, rel=1e-12)
    assert p.log_likelihood() == pytest.approx(full.log_likelihood(), rel=1e-11)
    for a, b in zip(p.gradient(), full.gradient()):
        assert rel_err(a, b) < 1e-10
    if ctor == "array":
        assert np.array_equal(p.ratings, R)
        with pytest.raises(ValueError):
            p.add_rating(*R[5])                          # already rated


def test_out_of_range_ids_are_an_error_not_a_fault(env):
    """ids outside the matrix are rejected when a list / pool is built (the reference asserts,
    pmf_cy.pyx:139-140); afterwards the device is still usable"""
    import torch
    from active_matrix_factorization_b200 import device as D, scoring as S
    i = np.array([0, 1, 5, 2], np.int32)
    j = np.array([0, 3, 1, 2], np.int32)
    r = np.ones(4)
    for bad_i, bad_j in ((np.array([0, 1, 6, 2], np.int32), j), (i, np.array([0, 4, 1, 2], np.int32)),
                         (np.array([0, -1, 5, 2], np.int32), j)):
        with pytest.raises(RuntimeError, match="outside 6 x 4"):
            D.Ratings(6, 4, bad_i, bad_j, r, "f64")
        with pytest.raises(RuntimeError, match="outside 6 x 4"):
            S.Pool(bad_i, bad_j, 6, 4, "f64", 3)
    rat = D.Ratings(6, 4, i, j, r, "f64")
    with pytest.raises(RuntimeError, match="amf_ratings_append"):
        rat.append(np.array([6], np.int32), np.array([0], np.int32), np.array([1.]))
    assert rat.nnz == 4
    rat.append(np.array([5], np.int32), np.array([3], np.int32), np.array([1.]))
    assert rat.nnz == 5
    rat.close()
    assert torch.zeros(4, device="cuda").sum().item() == 0      # no sticky CUDA error
