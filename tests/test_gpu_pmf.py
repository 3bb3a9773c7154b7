"""GPU parity of the PMF objective/gradient/fit path against the oracle and the reference's
golden outputs.  Every call goes through the C ABI (ctypes) of libamf_b200.so.

Tolerances: parity mode (f64) 1e-10 relative; fast mode (f32) 1e-5 relative to the scale of
the quantity (north_star: 1e-5 on objective and criteria)."""
import numpy as np
import pytest

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def amf():
    from active_matrix_factorization_b200 import build
    build.build()
    from active_matrix_factorization_b200 import pmf_cy
    return pmf_cy


def make_model(amf, R, U, V, d, dtype, subtract_mean=False, **hyp):
    p = amf.ProbabilisticMatrixFactorization(R, d, subtract_mean)
    p.compute_dtype = dtype
    p.users, p.items = U.copy(), V.copy()
    for k, v in hyp.items():
        setattr(p, k, v)
    return p


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-11), ("f32", 1e-5)])
@pytest.mark.parametrize("name", ["known_answer_10x10_d2", "random_12x20_d5"])
def test_ll_grad_golden(amf, golden, name, dtype, tol):
    g = golden(name)
    hyp = {k: float(g[k]) for k in ("sigma_sq", "sigma_u_sq", "sigma_v_sq") if k in g}
    p = make_model(amf, g["ratings"], g["users"], g["items"], g["users"].shape[1], dtype, **hyp)
    assert p.log_likelihood() == pytest.approx(float(g["ll"]), rel=tol)
    assert p.full_ll() == pytest.approx(float(g["full_ll"]), rel=tol)
    gu, gv = p.gradient()
    assert rel_err(gu, g["grad_u"]) < tol and rel_err(gv, g["grad_v"]) < tol
    # explicit-argument form used by fit_lls / check_grad
    assert p.log_likelihood(g["users"] * 1.01, g["items"] * .99) == pytest.approx(
        O.log_likelihood(g["ratings"], g["users"] * 1.01, g["items"] * .99, **hyp), rel=tol)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-5)])
@pytest.mark.parametrize("n,m,d,nnz,sm", [(300, 200, 32, 20000, False), (257, 129, 10, 5000, True),
                                           (64, 50, 1, 900, True), (500, 400, 15, 30000, False),
                                           (40, 30, 48, 700, False), (33, 21, 100, 400, True)])
def test_ll_grad_random(amf, n, m, d, nnz, sm, dtype, tol):
    rng = np.random.RandomState(n + d)
    cells = rng.permutation(n * m)[:nnz]
    ii, jj = cells // m, cells % m
    ii[0], jj[0] = n - 1, m - 1              # pin the matrix shape
    R = np.column_stack((ii, jj, rng.normal(3, 1, nnz)))
    U, V = rng.normal(0, .5, (n, d)), rng.normal(0, .5, (m, d))
    p = make_model(amf, R, U, V, d, dtype, sm, sigma_sq=.8, sigma_u_sq=7., sigma_v_sq=12.)
    h = dict(sigma_sq=.8, sigma_u_sq=7., sigma_v_sq=12., mean_rating=p.mean_rating, subtract_mean=sm)
    assert p.log_likelihood() == pytest.approx(O.log_likelihood(R, U, V, **h), rel=tol)
    gu, gv = p.gradient()
    ou, ov = O.gradient(R, U, V, **h)
    assert rel_err(gu, ou) < tol and rel_err(gv, ov) < tol
    # mini-batch form: gradient(ratings=batch)  (COO kernel with atomics)
    bu, bv = p.gradient(R[100:350])
    ou, ov = O.gradient(R[100:350], U, V, **h)
    assert rel_err(bu, ou) < tol and rel_err(bv, ov) < tol


def test_empty_rows_and_heavy_rows(amf):
    """ragged input: users/items with no ratings, one item rated by everyone, duplicates kept"""
    rng = np.random.RandomState(1)
    n, m, d = 70, 40, 8
    rows = [(i, 7, rng.normal()) for i in range(0, n, 1)]            # heavy column
    rows += [(3, j, rng.normal()) for j in range(0, m, 3) if j != 7]  # one heavy-ish row
    rows += [(n - 1, m - 1, 1.0), (n - 1, m - 1, 2.0)]                # duplicate cell
    R = np.array(rows)
    U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    for dtype, tol in (("f64", 1e-11), ("f32", 1e-5)):
        p = make_model(amf, R, U, V, d, dtype)
        assert p.log_likelihood() == pytest.approx(O.log_likelihood(R, U, V), rel=tol)
        gu, gv = p.gradient()
        ou, ov = O.gradient(R, U, V)
        assert rel_err(gu, ou) < tol and rel_err(gv, ov) < tol


@pytest.mark.parametrize("sm", [False, True])
def test_fit_lls_trajectory_golden(amf, golden, sm):
    """device-resident line search reproduces the reference's accepted-step sequence"""
    g = golden("fit_30x40_d4")
    tag = "_sm" if sm else ""
    p = make_model(amf, g["ratings"], g["users0"], g["items0"], 4, "f64", sm)
    lls = list(p.fit_lls())
    assert len(lls) == len(g["lls" + tag])
    np.testing.assert_allclose(lls, g["lls" + tag], rtol=1e-9)
    np.testing.assert_allclose(p.users, g["users_fit" + tag], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(p.items, g["items_fit" + tag], rtol=1e-6, atol=1e-8)
    assert p.rmse(g["real"]) == pytest.approx(float(g["rmse" + tag]), rel=1e-6)
    p.update_sigma(); p.update_sigma_uv()
    np.testing.assert_allclose([p.sigma_sq, p.sigma_u_sq, p.sigma_v_sq], g["sigmas" + tag], rtol=1e-7)


def test_fit_lls_fast_mode_objective(amf, golden):
    """f32 fast mode: final objective within 1e-5 relative of the fp64 reference fit"""
    g = golden("fit_30x40_d4")
    p = make_model(amf, g["ratings"], g["users0"], g["items0"], 4, "f32")
    lls = list(p.fit_lls())
    ref = g["lls"]
    assert abs(len(lls) - len(ref)) <= max(3, len(ref) // 10)
    k = min(len(lls), len(ref)) // 2
    np.testing.assert_allclose(lls[:k], ref[:k], rtol=1e-5)
    assert lls[-1] == pytest.approx(ref[-1], rel=1e-4, abs=2e-2)   # both stop within stop_thresh


def test_minibatch_sgd_matches_oracle_arithmetic(amf):
    rng = np.random.RandomState(2)
    n, m, d = 25, 30, 6
    R = np.column_stack((rng.randint(0, n, 400), rng.randint(0, m, 400), rng.normal(size=400)))
    R[0, :2] = (n - 1, m - 1)
    U, V = rng.normal(0, .3, (n, d)), rng.normal(0, .3, (m, d))
    p = make_model(amf, R.copy(), U, V, d, "f64")
    np.random.seed(7)
    it = p.fit_minibatches(64, lr=.5, momentum=.8)
    errs = [next(it) for _ in range(3)]
    # restatement of pmf_cy.pyx:308-351 with the oracle gradient
    np.random.seed(7)
    R2, U2, V2 = R.copy(), U.copy(), V.copy()
    ui, vi = np.zeros_like(U2), np.zeros_like(V2)
    mom, lr = float(np.float32(.8)), float(np.float32(.5))
    ref = []
    for _ in range(3):
        np.random.shuffle(R2)
        for s in range(0, 400, 64):
            b = R2[s:s + 64]
            gu, gv = O.gradient(b, U2, V2)
            ui = ui * mom + gu * (lr / len(b)); U2 = U2 + ui
            vi = vi * mom + gv * (lr / len(b)); V2 = V2 + vi
        ref.append(float(np.float32(np.sqrt(np.mean((O.predictions(R2, U2, V2) - R2[:, 2]) ** 2)))))
    np.testing.assert_allclose(errs, ref, rtol=1e-6)
    np.testing.assert_allclose(p.users, U2, rtol=1e-9, atol=1e-12)


def test_linearity_at_scale(amf):
    """size-independent property at a few million ratings: the data term of the gradient is
    linear in the ratings and the squared error matches a direct einsum"""
    import torch
    from active_matrix_factorization_b200 import device as D
    rng = np.random.RandomState(0)
    n, m, d, nnz = 20000, 5000, 32, 3_000_000
    ii, jj = rng.randint(0, n, nnz).astype(np.int32), rng.randint(0, m, nnz).astype(np.int32)
    r = rng.normal(size=nnz)
    U, V = rng.normal(0, .2, (n, d)), rng.normal(0, .2, (m, d))
    Ut, Vt = D.to_padded(U, "f32"), D.to_padded(V, "f32")
    prm = D.pmf_params(1., 10., 10., 0.)
    out = []
    for scale in (1.0, 2.0):
        rat = D.Ratings(n, m, ii, jj, r * scale, "f32")
        gu, gv = torch.empty_like(Ut), torch.empty_like(Vt)
        sums = D.loss_grad(rat, d, Ut, Vt, prm, gu, gv).cpu().numpy()
        out.append((sums, gu.double().cpu().numpy(), gv.double().cpu().numpy()))
        rat.close()
    pred = np.einsum("nd,nd->n", U[ii], V[jj])
    assert out[0][0][0] == pytest.approx(((r - pred) ** 2).sum(), rel=1e-5)
    assert out[0][0][1] == pytest.approx((U * U).sum(), rel=1e-5)
    # g(2r) - g(r) = sum_j V_j r / sigma^2  (the prior and prediction terms cancel)
    lin_u = np.zeros((n, d)); np.add.at(lin_u, ii, V[jj] * r[:, None])
    lin_v = np.zeros((m, d)); np.add.at(lin_v, jj, U[ii] * r[:, None])
    assert rel_err(out[1][1] - out[0][1], lin_u) < 1e-5
    assert rel_err(out[1][2] - out[0][2], lin_v) < 1e-5


def test_from_coo_matches_tuple_constructor(amf):
    """the COO constructor (no (nnz,3) array, no N*M sets) gives the same
    objective, gradient and fit as the reference-style constructor; device tensors accepted."""
    import torch
    rng = np.random.RandomState(4)
    n, m, d, nnz = 120, 90, 8, 3000
    cells = rng.permutation(n * m)[:nnz]
    ii, jj = cells // m, cells % m
    ii[0], jj[0] = n - 1, m - 1
    r = rng.normal(3, 1, nnz)
    R = np.column_stack((ii, jj, r)).astype(float)
    U, V = rng.normal(0, .3, (n, d)), rng.normal(0, .3, (m, d))
    a = make_model(amf, R, U, V, d, "f64", True)
    PMF = amf.ProbabilisticMatrixFactorization
    for args in ((ii, jj, r), (torch.from_numpy(ii).cuda(), torch.from_numpy(jj).cuda(), torch.from_numpy(r).cuda())):
        b = PMF.from_coo(*args, num_users=n, num_items=m, latent_d=d, subtract_mean=True,
                         init=(U.copy(), V.copy()))
        assert b.unrated == set() and b.mean_rating == pytest.approx(a.mean_rating, rel=1e-12)
        assert b.log_likelihood() == pytest.approx(a.log_likelihood(), rel=1e-12)
        assert b.full_ll() == pytest.approx(a.full_ll(), rel=1e-12)
        ga, gb = a.gradient(), b.gradient()
        np.testing.assert_allclose(gb[0], ga[0], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(gb[1], ga[1], rtol=1e-11, atol=1e-12)
        assert b.ratings.shape == (nnz, 3)                    # lazily materialised, same content
        assert sorted(map(tuple, b.ratings)) == sorted(map(tuple, R))
    b = PMF.from_coo(ii, jj, r, n, m, d, True, init=(U.copy(), V.copy()))
    a2 = make_model(amf, R, U, V, d, "f64", True)
    la, lb = list(a2.fit_lls()), list(b.fit_lls())
    assert len(la) == len(lb)
    np.testing.assert_allclose(lb, la, rtol=1e-9)
    np.testing.assert_allclose(b.users, a2.users, rtol=1e-7, atol=1e-9)


def test_from_coo_file_formats(amf, tmp_path):
    """SURVEY.md 8f-3 loaders on the device path: an i/j/r archive, the reference's `_ratings`
    dictionary (choose_training.py:215-259), a .npy table and text lines all give the model the
    tuple constructor builds"""
    import pickle
    rng = np.random.RandomState(6)
    n, m, d, nnz = 40, 30, 4, 300
    cells = rng.permutation(n * m)[:nnz]
    ii, jj = cells // m, cells % m
    ii[0], jj[0] = n - 1, m - 1
    r = rng.randint(1, 6, nnz).astype(float)
    R = np.column_stack((ii, jj, r)).astype(float)
    U, V = rng.normal(0, .3, (n, d)), rng.normal(0, .3, (m, d))
    want = make_model(amf, R, U, V, d, "f64", True)
    PMF = amf.ProbabilisticMatrixFactorization
    paths = []
    p = str(tmp_path / "a.npz"); np.savez(p, i=ii, j=jj, r=r, shape=np.array([n, m])); paths.append(p)
    p = str(tmp_path / "b.npz"); np.savez(p, _ratings=R, _real=np.zeros((n, m))); paths.append(p)
    p = str(tmp_path / "c.pkl")
    with open(p, "wb") as f:
        pickle.dump({"_ratings": R, "_real": np.zeros((n, m)), "_rating_vals": (1, 2, 3, 4, 5)}, f)
    paths.append(p)
    p = str(tmp_path / "d.npy"); np.save(p, R); paths.append(p)
    p = str(tmp_path / "e.txt"); np.savetxt(p, R); paths.append(p)
    for p in paths:
        b = PMF.from_coo_file(p, d, True, init=(U.copy(), V.copy()))
        assert (b.num_users, b.num_items) == (n, m), p
        assert b.log_likelihood() == pytest.approx(want.log_likelihood(), rel=1e-12), p
        np.testing.assert_allclose(b.gradient()[0], want.gradient()[0], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("sm", [False, True])
def test_device_fit_matches_reference_trajectory(amf, golden, sm):
    """amf_pmf_fit_lls: the whole line search in one launch takes the reference's accepted steps
    (same count, objectives to 1e-9) and ends at the reference's factors; fit() uses it"""
    import ctypes as C
    import torch
    from active_matrix_factorization_b200 import _native as N, device as D
    g = golden("fit_30x40_d4")
    tag = "_sm" if sm else ""
    p = make_model(amf, g["ratings"], g["users0"], g["items0"], 4, "f64", sm)
    lib = N.require_device()
    rat = p._rating_handle()
    ld = D.padded_ld(4, "f64")
    U, V = D.to_padded(g["users0"], "f64"), D.to_padded(g["items0"], "f64")
    nbytes = int(lib.amf_pmf_fit_workspace_bytes(rat.handle, N.F64, ld))
    work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    trace = torch.zeros(4096, dtype=torch.float64, device="cuda")
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    params = p._params()
    N.check(lib.amf_pmf_fit_lls(rat.handle, N.F64, 4, ld, D.ptr(U), D.ptr(V), C.byref(params),
                                p.learning_rate, p.min_learning_rate, p.stop_thresh, 0, D.ptr(trace),
                                4096, D.ptr(res), D.ptr(work), nbytes, D.stream_ptr()))
    steps = int(res[2:3].view(torch.int32)[0].item())
    ref = g["lls" + tag]
    assert steps == len(ref)
    np.testing.assert_allclose(trace[:steps].cpu().numpy(), ref, rtol=1e-9)
    np.testing.assert_allclose(D.from_padded(U, 4), g["users_fit" + tag], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(D.from_padded(V, 4), g["items_fit" + tag], rtol=1e-6, atol=1e-8)
    assert res[1].item() == pytest.approx(ref[-1], rel=1e-9)
    # the class method takes the same path
    p.fit()
    np.testing.assert_allclose(p.users, g["users_fit" + tag], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(p.items, g["items_fit" + tag], rtol=1e-6, atol=1e-8)
    # a step cap stops early with the same prefix
    U2, V2 = D.to_padded(g["users0"], "f64"), D.to_padded(g["items0"], "f64")
    N.check(lib.amf_pmf_fit_lls(rat.handle, N.F64, 4, ld, D.ptr(U2), D.ptr(V2), C.byref(params),
                                p.learning_rate, p.min_learning_rate, p.stop_thresh, 5, D.ptr(trace),
                                4096, D.ptr(res), D.ptr(work), nbytes, D.stream_ptr()))
    assert int(res[2:3].view(torch.int32)[0].item()) == 5
    np.testing.assert_allclose(trace[:5].cpu().numpy(), ref[:5], rtol=1e-9)


def test_device_fit_fast_mode(amf, golden):
    """f32: the one-launch fit ends within stop_thresh of the fp64 reference objective"""
    g = golden("fit_30x40_d4")
    p = make_model(amf, g["ratings"], g["users0"], g["items0"], 4, "f32")
    p.fit()
    assert p.log_likelihood() == pytest.approx(float(g["lls"][-1]), rel=1e-4, abs=2e-2)
