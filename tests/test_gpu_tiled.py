"""GPU parity of the shared-memory-tiled loss+gradient (csrc/tiled.cu) against the oracle and
against the row-sorted kernels, through the C ABI.  The tiled layout is forced with
amf_ratings_set_layout so that small and ragged inputs exercise it (AUTO only picks it from
2^20 ratings up).

Tolerances: parity mode (f64) 1e-10 relative; fast mode (f32) 1e-5 relative to the scale of the
quantity (north_star: 1e-5 on objective and criteria)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from active_matrix_factorization_b200 import build
    build.build()
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    N.require_device()
    return N, D, torch


def loss_grad(env, rat, U, V, d, ld, dtype, hyp, mean_offset=0.0, grad=True):
    N, D, torch = env
    lib = N.require_device()
    dt = D.torch_dtype(dtype)
    Ut = torch.zeros((U.shape[0], ld), dtype=dt, device="cuda"); Ut[:, :d] = torch.from_numpy(U).to(dt)
    Vt = torch.zeros((V.shape[0], ld), dtype=dt, device="cuda"); Vt[:, :d] = torch.from_numpy(V).to(dt)
    dU, dV = (torch.full_like(Ut, 7.0), torch.full_like(Vt, 7.0)) if grad else (None, None)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    params = D.pmf_params(hyp["sigma_sq"], hyp["sigma_u_sq"], hyp["sigma_v_sq"], mean_offset)
    N.check(lib.amf_pmf_loss_grad(rat.handle, D.code(dtype), d, ld, D.ptr(Ut), D.ptr(Vt),
                                  C.byref(params), D.ptr(dU), D.ptr(dV), D.ptr(sums), D.stream_ptr()))
    s = sums.cpu().numpy()
    ll = -s[0] / (2 * hyp["sigma_sq"]) - s[1] / (2 * hyp["sigma_u_sq"]) - s[2] / (2 * hyp["sigma_v_sq"])
    if not grad:
        return ll, None, None
    return ll, dU[:, :d].double().cpu().numpy(), dV[:, :d].double().cpu().numpy()


def rel_err(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


HYP = dict(sigma_sq=.8, sigma_u_sq=7., sigma_v_sq=12.)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
@pytest.mark.parametrize("n,m,d,nnz", [(300, 200, 32, 20000),      # one tile per side
                                        (5000, 3000, 32, 200000),   # several tiles per side
                                        (700, 2500, 16, 60000),
                                        (90, 4000, 64, 30000),      # f64: 512-byte rows -> unsupported
                                        (64, 50, 8, 900)])
def test_tiled_matches_oracle_and_rows(env, n, m, d, nnz, dtype, tol):
    N, D, torch = env
    rng = np.random.RandomState(n + d)
    cells = rng.permutation(n * m)[:nnz]
    ii, jj = (cells // m).astype(np.int32), (cells % m).astype(np.int32)
    r = rng.normal(3, 1, nnz)
    R = np.column_stack((ii, jj, r))
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    rat = D.Ratings(n, m, ii, jj, r, dtype)
    ld = D.padded_ld(d, dtype)
    row_bytes = ld * (4 if dtype == "f32" else 8)
    rat.set_layout("rows")
    ll_r, gu_r, gv_r = loss_grad(env, rat, U, V, d, ld, dtype, HYP, mean_offset=.25)
    rat.set_layout("tiled")
    if row_bytes not in (64, 128, 256):
        with pytest.raises(RuntimeError, match="tiled rating list"):
            loss_grad(env, rat, U, V, d, ld, dtype, HYP, mean_offset=.25)
        return
    ll_t, gu_t, gv_t = loss_grad(env, rat, U, V, d, ld, dtype, HYP, mean_offset=.25)
    ll_only, _, _ = loss_grad(env, rat, U, V, d, ld, dtype, HYP, mean_offset=.25, grad=False)
    h = dict(HYP, mean_rating=.25, subtract_mean=True)
    ll_o = O.log_likelihood(R, U, V, **h)
    gu_o, gv_o = O.gradient(R, U, V, **h)
    assert ll_t == pytest.approx(ll_o, rel=tol) and ll_only == pytest.approx(ll_o, rel=tol)
    assert rel_err(gu_t, gu_o) < tol and rel_err(gv_t, gv_o) < tol
    assert ll_t == pytest.approx(ll_r, rel=tol)
    assert rel_err(gu_t, gu_r) < tol and rel_err(gv_t, gv_r) < tol


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
def test_tiled_ragged(env, dtype, tol, monkeypatch):
    """users/items with no ratings, one item rated by everyone and one heavy user (runs far longer
    than a 64-entry segment), duplicate cells, a last tile with a single row"""
    N, D, torch = env
    rng = np.random.RandomState(1)
    monkeypatch.setenv("AMF_TILED_KB", "64")   # 512-row (fp32) / 256-row (fp64) tiles
    n, m, d = 1537, 3585, 32          # both 1 mod 512: the last tile of either side holds 1 row
    rows = [(i, 7, rng.normal()) for i in range(n)]
    rows += [(3, j, rng.normal()) for j in range(0, m, 3) if j != 7]
    rows += [(n - 1, m - 1, 1.0), (n - 1, m - 1, 2.0), (0, m - 1, -1.0)]
    R = np.array(rows, float)
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    rat = D.Ratings(n, m, R[:, 0].astype(np.int32), R[:, 1].astype(np.int32), R[:, 2], dtype)
    rat.set_layout("tiled")
    ld = D.padded_ld(d, dtype)
    ll, gu, gv = loss_grad(env, rat, U, V, d, ld, dtype, HYP)
    assert ll == pytest.approx(O.log_likelihood(R, U, V, **HYP), rel=tol)
    ou, ov = O.gradient(R, U, V, **HYP)
    assert rel_err(gu, ou) < tol and rel_err(gv, ov) < tol
    # rows without ratings carry the prior term only
    empty_items = np.setdiff1d(np.arange(m), R[:, 1].astype(int))
    np.testing.assert_allclose(gv[empty_items], -V[empty_items] / HYP["sigma_v_sq"], rtol=tol)


def test_tiled_more_tiles_than_threads(env, monkeypatch):
    """200,000 items in 128-row tiles = 1563 item tiles, more than the CTA's 384 threads: the
    search for the next tile with work left takes several rounds; result equal to the row-sorted
    kernels and the oracle"""
    N, D, torch = env
    monkeypatch.setenv("AMF_TILED_KB", "16")
    rng = np.random.RandomState(3)
    n, m, d, nnz = 300, 200_000, 32, 100_000
    ii = rng.randint(0, n, nnz).astype(np.int32)
    jj = rng.randint(0, m, nnz).astype(np.int32)
    r = rng.normal(3, 1, nnz)
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    rat = D.Ratings(n, m, ii, jj, r, "f32")
    ld = D.padded_ld(d, "f32")
    rat.set_layout("rows")
    ll_r, gu_r, gv_r = loss_grad(env, rat, U, V, d, ld, "f32", HYP)
    rat.set_layout("tiled")
    ll_t, gu_t, gv_t = loss_grad(env, rat, U, V, d, ld, "f32", HYP)
    assert ll_t == pytest.approx(ll_r, rel=2e-5)
    assert rel_err(gu_t, gu_r) < 2e-5 and rel_err(gv_t, gv_r) < 2e-5
    R = np.column_stack((ii, jj, r))
    assert ll_t == pytest.approx(O.log_likelihood(R, U, V, **HYP), rel=2e-5)


def test_tiled_auto_threshold_and_model_api(env):
    """AUTO switches to the tiled copy from 2^20 ratings; the drop-in class sees the same
    objective and gradient either way (f32, 1e-5)"""
    N, D, torch = env
    from active_matrix_factorization_b200.pmf_cy import ProbabilisticMatrixFactorization as PMF
    rng = np.random.RandomState(2)
    n, m, d, nnz = 4000, 3000, 32, (1 << 20) + 5000
    cells = rng.permutation(n * m)[:nnz]
    R = np.column_stack((cells // m, cells % m, rng.normal(3, 1, nnz))).astype(float)
    U, V = rng.normal(0, .3, (n, d)), rng.normal(0, .3, (m, d))
    out = {}
    for layout in ("rows", "auto"):
        p = PMF.from_coo(R[:, 0].astype(np.int32), R[:, 1].astype(np.int32), R[:, 2], n, m, d,
                         init=(U.copy(), V.copy()))
        p.compute_dtype = "f32"
        p._rating_handle().set_layout(layout)
        out[layout] = (p.log_likelihood(), p.gradient())
    assert out["auto"][0] == pytest.approx(out["rows"][0], rel=1e-5)
    for a, b in zip(out["auto"][1], out["rows"][1]):
        assert rel_err(a, b) < 2e-5


@pytest.mark.parametrize("layout", ["rows", "tiled"])
def test_two_part_pass_equals_one_call(env, layout):
    """amf_pmf_loss_grad_part: part 0 (priors, dU, sums) then part 1 (dV), with a capped grid,
    gives what the single call gives"""
    N, D, torch = env
    lib = N.require_device()
    rng = np.random.RandomState(21)
    n, m, d, nnz = 3000, 2500, 32, 150000
    cells = rng.permutation(n * m)[:nnz]
    ii, jj, r = (cells // m).astype(np.int32), (cells % m).astype(np.int32), rng.normal(3, 1, nnz)
    rat = D.Ratings(n, m, ii, jj, r, "f32")
    rat.set_layout(layout)
    U = D.to_padded(rng.normal(0, .5, (n, d)), "f32")
    V = D.to_padded(rng.normal(0, .5, (m, d)), "f32")
    params = D.pmf_params(.8, 7., 12., .25)
    dU, dV = torch.empty_like(U), torch.empty_like(V)
    sums = D.loss_grad(rat, d, U, V, params, dU, dV)
    gU, gV = torch.full_like(U, 9.), torch.full_like(V, 9.)
    s2 = torch.full((3,), 5., dtype=torch.float64, device="cuda")
    D.loss_grad_part(rat, d, U, V, params, gU, gV, s2, 0)
    # after part 0: dU and the sums are final, dV holds the prior term only
    assert rel_err(gU.cpu().numpy(), dU.cpu().numpy()) < 2e-6
    assert torch.allclose(s2, sums, rtol=1e-9)
    assert rel_err(gV.cpu().numpy(), (-V / 12.).cpu().numpy()) < 1e-6
    D.loss_grad_part(rat, d, U, V, params, gU, gV, s2, 1, max_ctas=5)
    assert rel_err(gV.cpu().numpy(), dV.cpu().numpy()) < 2e-6
    assert rel_err(gU.cpu().numpy(), dU.cpu().numpy()) < 2e-6
